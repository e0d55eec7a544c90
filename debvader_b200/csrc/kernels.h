// Internal (non-ABI) declarations shared between the .cu files of libdebvader_b200.
#pragma once
#include "epilogue.cuh"
#include <cuda.h>

namespace dbv {

// ---- fp32 SIMT gather convolution (simt_kernels.cu) ----------------------------------------------
struct SimtConv {
  const float* in;  // NHWC fp32 [B][Hin][Win][Cin]
  const float* w;   // gather form [ksz*ksz][Cin][CoutP]
  long long B;
  int Hin, Win, Cin, Hout, Wout, CoutP;
  int mode;         // 0: iy = stride*y + ky - pb    1: stride-2 transposed conv (iy = (y-ky)/2)
  int stride, pb, ksz;
  const float* in_scale;  // BatchNorm folded to scale/shift, applied to in-bounds pixels only
  const float* in_shift;
  OutSpec o;
};
// Launch with programmatic dependent launch when enabled (DBV_PDL, see tc_ptx.cuh:pdl_wait): only for kernels that call pdl_wait()
// before their first access to anything another kernel of the stream produces or consumes.
bool pdl_enabled();
template <typename Kernel, typename Arg>
static inline cudaError_t launch_pdl(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t st, const Arg& arg) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}

int launch_simt_conv(const SimtConv& p, cudaStream_t st);
int launch_latent(const float* params, const float* eps, unsigned long long seed, int sample, long long first_stamp,
                  long long B, float* z, float* loc, float* std_out, float* zp, const float* alpha0, cudaStream_t st);
int launch_prelu_vec(const float* z, const float* alpha, long long n, int C, float* out, cudaStream_t st);
int launch_cast_f64_f32(const double* in, float* out, long long n, cudaStream_t st);
int launch_act_to_f32(const OutSpec& o, long long B, float* out, cudaStream_t st);
int launch_bn_pack8(const float* x, const float* bn_scale, const float* bn_shift, long long B, const OutSpec& o, cudaStream_t st);

// ---- tcgen05 implicit-GEMM convolution (tc_conv.cu) ------------------------------------------------
// One k-block = one (tap, channel chunk): the A box(es) of the activation tensor (hi[, lo] plane) and the
// B box(es) of the packed weights (hi[, lo] part), multiplied by CBK/16 x {1, 2 or 3} tcgen05.mma of K=16.
struct TcKBlock {
  int16_t dx, dy;     // offset of the A box relative to the output tile origin (pixels)
  int16_t plane;      // parity plane of the input tensor (0 for plain tensors)
  int16_t c_off;      // first channel of the A box (hi plane: [0,Cin), lo plane: [Cpad,Cpad+Cin))
  int32_t b_row;      // first row of the B box in the packed weight tensor
};

constexpr int TC_MAX_KB = 192;  // enc_dense in bf16x3: 16 pixels x 4 chunks x 3 pairings
constexpr int TC_MAX_CLS = 4;

struct TcClass {
  int kb_begin, nkb;  // slice of the k-block table
  int oy0, ox0;       // output pixel = (oy0 + osy*ty, ox0 + osx*tx) for tile-space pixel (ty,tx)
  int osy, osx;
};

struct TcLayer {
  CUtensorMap tmA;    // 5D (C, W, H, P, B) bf16
  CUtensorMap tmB;    // 2D (CBK, rows) bf16 packed weights
  TcKBlock kb[TC_MAX_KB];
  TcClass cls[TC_MAX_CLS];
  int n_cls;
  int TW, TH, TB;     // tile box in tile space (TW*TH*TB <= 128 rows)
  int SH, SW;         // tile-space image extents (valid outputs)
  int tiles_x, tiles_y;
  int n_tiles_n;      // N tiling (dense layers); channel offset = nt * NT
  int nt_pixel_mode;  // Dense -> Reshape(4,4,256): 0 = off, else N tiles per output pixel (256 / NT): tile nt covers channels
                      // (nt % m) * NT.. of pixel nt / m, bias index = flat (pixel, channel)
  int seg_kb;         // > 0 (DBV_PREC_FP32TC): k-blocks chained per TMEM accumulator before the epilogue promotes the partial sum
  long long B;        // stamps in this launch
  long long tiles_per_cls;
  long long total_tiles;
  int a_bytes, b_bytes;  // bytes of one activation box / one weight box (expect_tx = parts * (a_bytes + b_bytes))
  int ab_f16;            // operand format of this layer's activations and weights: 0 = bf16, 1 = fp16
  int x3;                // hi/lo split precision: a k-block loads A_hi, A_lo, B_hi, B_lo once and issues all three pairings
  int lo_coff;           // channel offset of the lo plane in the activation tensor (input Cpad)
  int lo_brow;           // row offset of the lo weight block relative to the hi block (Ntot)
  int wide;              // A_hi x [B_hi | B_lo] as one MMA of N = 2*NT (set by tc_stage_plan when 2*NT <= 256)
  int stages, stage_bytes;
  long long pair_items;  // CTA-pair kernel: n_cls x ceil(m tiles / 2) x n_tiles_n cluster work items
  int dbg_shift_rows;    // probe only: A box loaded dbg_shift_rows batch rows early, descriptor start advanced to compensate
  int dbg_base_mode;     // probe only: 0 = base_offset field 0, 1 = (start_addr >> 7) & 7
  OutSpec o;
};

// CBK: K elements per k-block (32 -> 64-byte rows / SWIZZLE_64B, 64 -> 128-byte rows / SWIZZLE_128B)
// NT : MMA N (multiple of 16, 16..256)
int launch_tc_layer(const TcLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st);
void tc_stage_plan(TcLayer& L, int CBK, int NT);  // call after x3 is set

// ---- the same GEMM on CTA pairs (tc_pair.cu): cta_group::2, M = 256, each CTA holds half of every weight box ----
// TcLayer.tmB must have a box of NT/2 rows.
int launch_tc_pair(const TcLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st);
bool tc_pair_supported(int CBK, int NT);
void tc_pair_stage_plan(TcLayer& L, int CBK, int NT);
bool tc_layer_supported(int CBK, int NT);
bool tc_seg_supported(int CBK, int NT);

// ---- CTA-pair GEMM with a per-chunk halo box (tc_pairh.cu): 8x8 .. 16x16 maps, N = 128 / 256 -----------------------------
struct PairHLayer {
  CUtensorMap tmA;   // (C, W, B, H, 1) bf16, box (64, 10, TB, HB, 1), SWIZZLE_128B: shared memory [row][stamp][10 px][64 ch]
  CUtensorMap tmB;   // packed weights, box (64, NT/2)
  TcClass cls[TC_MAX_CLS];
  int n_cls;
  int ab_f16;        // operand format: 0 = bf16, 1 = fp16
  int cls_begin[TC_MAX_CLS + 1];  // taps of class c: [cls_begin[c], cls_begin[c+1])
  int cls_groups;    // 1, or 2 when the accumulators of all classes do not fit 256 TMEM columns: a work item then covers one
                     // half of the classes (the halo boxes are loaded once per half) and the accumulators stay double buffered
  int tap_aoff[16];  // byte offset of the tap's window in the halo box: ((dy+1) * TB * 10 + dx+1) * 128
  int tap_brow[16];  // first weight row of the tap's hi block for chunk 0
  int nchunk, chunk_brow, lo_brow, lo_coff;
  int TB, tiles_x, SW, SH;
  int abox_bytes;    // one halo box rounded up to 1024
  int abox_tx;       // bytes one box load delivers
  int a_stages, b_stages, tail_pad, smem_bytes;
  long long B, pair_items;
  OutSpec o;
};
int launch_tc_pairh(const PairHLayer& L, int NT, int max_ctas, cudaStream_t st);
bool tc_pairh_supported(int NT);

// ---- tcgen05 convolution with a resident halo tile (tc_halo.cu) ----------------------------------------
constexpr int HALO_MAX_SMEM = 232448;  // 227 KB
// One tcgen05.mma (K = 16) of a halo band.  The whole band is a FLAT, host-made list of these, cut into units by
// unit_op_end[]: the issuing thread reads an entry with one uniform 128-bit constant load, adds the two smem bases
// and issues — no per-tile / per-class index arithmetic (measured in round 2 with clock64: the old per-sub-unit loop
// set-up cost ~300 cycles and the per-MMA descriptor selection ~20, all of it tensor-pipe idle time because the MMA
// queue is shallow).
struct HaloOp {
  uint32_t a;      // A descriptor start (bytes >> 4) inside one halo buffer: region + tap shift + k-step + tile
  uint32_t b;      // B descriptor start (bytes >> 4) inside the resident weights
  uint32_t idesc;  // instruction descriptor (M = 128, this op's N, operand formats)
  uint32_t d;      // bits 0-15: first accumulator column inside the unit's slot; bit 16: accumulate (0 = overwrite)
};
constexpr int HALO_MAX_OPS = 288;    // MMAs per band
constexpr int HALO_MAX_UNITS = 64;   // units (one accumulator slot, one commit) per band
constexpr int HALO_MAX_ITEMS = 256;  // epilogue items (128 positions x NV channels of one class) per band
constexpr int HALO_MAX_WBLK = 96;    // resident weight blocks
struct HaloLayer {
  CUtensorMap tmA;  // 5D (C, W, H, 1, B) bf16, box (CBK, W+2, R+2, 1, 1)
  CUtensorMap tmB;  // packed weights, box (CBK, NT)
  float bias_c[64];        // the layer's bias (zero padded to NT) in the kernel's constant bank: added without a load
  HaloOp ops[HALO_MAX_OPS];
  uint16_t unit_op_end[HALO_MAX_UNITS];    // ops of unit k of a band: [unit_op_end[k-1], unit_op_end[k])
  uint16_t unit_item_end[HALO_MAX_UNITS];  // epilogue items of unit k: [unit_item_end[k-1], unit_item_end[k])
  uint32_t items[HALO_MAX_ITEMS];  // bits 0-8: first accumulator column of the item inside the slot; 9-10: output class;
                                   // 11-15: tile of the band; 16-18: channel chunk (c0 = chunk * NV)
  uint8_t w_src[HALO_MAX_WBLK];    // resident weight block j is block w_src[j] of the packed weight tensor
  TcClass cls[TC_MAX_CLS];
  int n_cls;
  int n_units;       // units per band
  int slot_cols;     // accumulator columns of a slot (power of two <= 256); the ring has 512 / slot_cols slots (<= 8)
  int ab_f16;        // operand format of this layer's activations and weights: 0 = bf16, 1 = fp16
  int W, H;          // tile-space extents (valid outputs sx < W, sy < H)
  int R, WP;         // output rows per band, W + 2*pad
  int pad;           // zero columns left of a row (slot 0 = x = -1)
  int pad_top;       // halo rows above the band (0 or 1: what the taps reach); the box holds R + pad_top + pad_bottom rows
  int ntiles;        // ceil(R * WP / 128)
  int n_regions;     // planes * channel chunks
  int region_coff[8];  // TMA coordinate 0 (first channel) of each region
  int region_c3[8];    // TMA coordinate 3: parity plane of a stride-2 Conv2D's input, else 0
  int region_bytes;  // (R+2) * WP * ROWB rounded up to 1024 (CG8: exact, the regions of a buffer are one contiguous TMA box)
  int buf_bytes;     // stride between halo buffers: n_regions * region_bytes (CG8: rounded up, with >= 16 zeroed bytes of slack)
  int a_box_bytes;   // (R+2) * WP * ROWB
  int n_wblk, w_rows_per_blk, w_bytes;
  const void* w_img;  // no-swizzle mode (conv1): the resident weights as one ready-made shared-memory image (bulk copy)
  uint32_t magic_wp;  // ceil(2^32 / WP): n / WP == __umulhi(n, magic) for n, WP < 2^16 (epilogue coordinates)
  int cg8;           // the input is OUT_BF16_CG8: ONE un-swizzled TMA box (u64 tensor map (2W, H, planes*Cin/8, B, 1)) fills all
                     // group-plane regions; the A operand is un-swizzled with LBO = region_bytes (K=16 = two groups)
  int pair;          // 1: CTA-pair plan (tc_halo2.cu): M = 256 ops, each CTA keeps one weight part (block w_src[j] + cluster rank),
                     // total_bands counts (band, stamp pair) items
  int nbuf;          // halo buffers in the ring (1 or 2)
  int U;             // sub-units per unit (plan parameter, kept for the logs)
  int nseg;          // DBV_PREC_FP32TC: a (class, tile)'s taps are spread over nseg accumulators seg_cols columns apart (shorter
  int seg_cols;      // tcgen05 accumulation chains); the epilogue sums them with round-to-nearest adds.  1 otherwise
  int wide;          // hi/lo weights: a class tile's accumulator = [A*B_hi (+ A_lo*B_hi) | A*B_lo] (2*NT columns, summed by the
                     // epilogue)
  int tail_pad;      // readable slack after the last halo buffer (garbage positions over-read < 129 rows)
  int smem_bytes;
  int dbg_skip;      // timing ablations only (env DBV_HALO_SKIP): bit0 skip the MMAs, bit1 skip the epilogue body
  int dbg_id;        // layer index for the clock64 counters of the ablation build
  int bands_per_img;
  long long B, total_bands;
  OutSpec o;
};
int launch_halo_layer(const HaloLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st);
// the same plan on CTA pairs (stride-1 layers with single-plane activations and hi/lo weights: the fp16 tail of DBV_PREC_MIXED)
int launch_halo_pair_layer(const HaloLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st);
bool halo_pair_supported(int CBK, int NT);
bool halo_layer_supported(int CBK, int NT);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes, int elem_bytes = 2);

}  // namespace dbv
