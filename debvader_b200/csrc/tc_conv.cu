// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// One persistent, warp-specialised kernel runs every GEMM-shaped layer of the network
// (Conv2D s1/s2, Conv2DTranspose s1/s2 as 4 output-parity classes, Dense):
//
//   D[128 pixels x NT channels] (fp32, TMEM)  +=  A[128 x CBK] (bf16, smem, K-major, swizzled)
//                                               x B[NT  x CBK] (bf16, smem, K-major, swizzled)
//
//   warp 0   TMA producer: per k-block one 5-D box of the NHWC activation tensor, shifted by the
//            tap offset (out-of-bounds elements are zero-filled by TMA = TF "SAME" padding), and
//            one 2-D box of the packed weights; both land on the stage's `full` mbarrier.
//   warp 1   MMA issuer: one elected lane issues CBK/16 tcgen05.mma (M=128, N=NT, K=16) per
//            k-block into a double-buffered TMEM accumulator; tcgen05.commit releases the smem
//            stage (`empty`) and, after the last k-block, publishes the accumulator (`tfull`).
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, bias + per-(h,w,c) PReLU (+ second
//            PReLU / ReLU), bf16 (hi[/lo]) conversion, 16-byte stores in the layout the next
//            layer's TMA expects; then `tempty` hands the accumulator back.
//
// Reference semantics implemented: model/model.py:80-98 (encoder convs + Dense),
// :117-137 (decoder Dense + Conv2DTranspose stack + head), with TF padding rules.
#include "tc_ptx.cuh"
#include <mutex>

namespace dbv {

constexpr int TC_THREADS = 64 + 2 * 128;  // TMA warp, MMA warp, two epilogue groups of 4 warps
constexpr int TC_NSLOT_MAX = 4;            // accumulator slots in the TMEM ring

template <int CBK, int NT>
struct TcCfg {
  static constexpr int ROWB = CBK * 2;
  static constexpr int A_BYTES = 128 * ROWB;  // one activation box (one plane)
  static constexpr int B_BYTES = NT * ROWB;   // one weight box (one part)
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b format [7,10)/[10,13),
  // K-major both, N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t IDESC =
      (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);  // + idesc_ab_fmt()
  static constexpr uint32_t IDESC2 =
      (1u << 4) | ((uint32_t)(((2 * NT) & 0x1ff) >> 3) << 17) | ((128u >> 4) << 24);
};

// smem stage of one k-block = (tap, channel chunk):  [A_hi | A_lo | B_hi | B_lo]  (the lo boxes only in the hi/lo split
// precisions).  Every box is loaded ONCE and used by all pairings:
//   wide (2*NT <= 256):  D[:, 0:2NT] += A_hi x [B_hi | B_lo]   (one MMA of N = 2*NT: B_hi, B_lo are adjacent)
//                        D[:, 0:NT ] += A_lo x B_hi            (the epilogue adds the two halves)
//   else              :  D += A_hi x B_hi;  D += A_hi x B_lo;  D += A_lo x B_hi
// Shared-memory bandwidth (TMA fill + MMA operand fetch) is what bounds these layers, so not re-loading A_hi and
// B_hi per pairing is worth 1.5x on the fill side.
// SEG (DBV_PREC_FP32TC): tcgen05 accumulates in fp32 WITHOUT rounding to nearest, so a long accumulation chain loses ~1 ulp per
// MMA with a bias (measured, tools/tc_accum_probe.cu: 5.5e-6 of the output scale at K = 2304 against 1.3e-6 for an fp32 FMA
// chain).  In this mode a tile's k-blocks are cut into segments of L.seg_kb k-blocks (<= 128 values of K); each segment
// accumulates into its own slot of the TMEM ring and the epilogue warps PROMOTE it: tcgen05.ld + add.rn.f32 into register
// accumulators (7.4e-7 at K = 2304).  Both epilogue groups work on every tile, each on alternate NV-column chunks, so a
// thread holds at most NT / 2 accumulators.
template <int CBK, int NT, bool SEG>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ TcLayer L) {
  pdl_trigger();
  using Cfg = TcCfg<CBK, NT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int STAGES = L.stages;
  const uint32_t stage_bytes = (uint32_t)L.stage_bytes;
  const uint32_t offB = (uint32_t)(L.x3 ? 2 : 1) * Cfg::A_BYTES;  // B_hi inside a stage; A_lo at A_BYTES, B_lo at offB + B_BYTES
  const uint32_t sBar = base + (uint32_t)STAGES * stage_bytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 64;
  const uint32_t bar_tfull = sBar + 128, bar_tempty = bar_tfull + 8 * TC_NSLOT_MAX;
  const uint32_t s_tmem = bar_tempty + 8 * TC_NSLOT_MAX;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t SLOTW = (uint32_t)tmem_cols_for(NT);  // slot pitch in columns (power of two >= NT)
  const uint32_t slot_pitch = L.wide ? 2 * SLOTW : SLOTW;
  const uint32_t nslot = (512u / slot_pitch) < (uint32_t)TC_NSLOT_MAX ? (512u / slot_pitch) : (uint32_t)TC_NSLOT_MAX;  // 2 or 4
  const uint32_t slot_shift = 31u - (uint32_t)__clz((int)nslot);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    tma_prefetch_desc(&L.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < TC_NSLOT_MAX; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, SEG ? 8 : 4);  // SEG: both epilogue groups drain every segment
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // nothing produced or consumed by the previous kernel is touched above this line

  const long long total = L.total_tiles;
  const int tiles_img = L.tiles_x * L.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)((L.x3 ? 2 : 1) * (L.a_bytes + L.b_bytes));
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int c = (int)(t / L.tiles_per_cls);
        long long r = t - (long long)c * L.tiles_per_cls;
        const int nt = (int)(r % L.n_tiles_n);
        r /= L.n_tiles_n;
        const int ti = (int)(r % tiles_img);
        const int bt = (int)(r / tiles_img);
        const int x0 = (ti % L.tiles_x) * L.TW, y0 = (ti / L.tiles_x) * L.TH, b0 = bt * L.TB;
        const TcClass cl = L.cls[c];
        for (int kb = 0; kb < cl.nkb; ++kb) {
          const TcKBlock K = L.kb[cl.kb_begin + kb];
          const uint32_t sS = base + (uint32_t)stage * stage_bytes, bar = bar_full + 8 * stage;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          mbar_expect_tx(bar, tx);
          tma_load_5d(sS, &L.tmA, bar, K.c_off, x0 + K.dx, y0 + K.dy, K.plane, b0 - DBV_DBG(L.dbg_shift_rows));
          tma_load_2d(sS + offB, &L.tmB, bar, 0, K.b_row + nt * NT);
          if (L.x3) {
            tma_load_5d(sS + Cfg::A_BYTES, &L.tmA, bar, K.c_off + L.lo_coff, x0 + K.dx, y0 + K.dy, K.plane, b0 - DBV_DBG(L.dbg_shift_rows));
            tma_load_2d(sS + offB + Cfg::B_BYTES, &L.tmB, bar, 0, K.b_row + L.lo_brow + nt * NT);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;
      constexpr uint32_t HI = smem_desc_hi<Cfg::ROWB>();
      const uint32_t IDESC = Cfg::IDESC | idesc_ab_fmt(L.ab_f16), IDESC2 = Cfg::IDESC2 | idesc_ab_fmt(L.ab_f16);
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int c = (int)(t / L.tiles_per_cls);
        const int nkb = L.cls[c].nkb;
        const int seg = SEG ? L.seg_kb : nkb;  // k-blocks chained into one accumulator
        for (int kb0 = 0; kb0 < nkb; kb0 += seg, ++u) {
          const uint32_t slot = u & (nslot - 1);
          mbar_wait(bar_tempty + 8 * slot, ((u >> slot_shift) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * slot_pitch;
          const int kb1 = kb0 + seg < nkb ? kb0 + seg : nkb;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sS = base + (uint32_t)stage * stage_bytes;
            uint32_t a_addr = sS + DBV_DBG(L.dbg_shift_rows) * Cfg::ROWB;
            uint32_t ahi = kSmemDescLoConst | ((a_addr & 0x3FFFFu) >> 4);
            uint32_t HIA = HI;
            if (DBV_DBG(L.dbg_base_mode) == 1) HIA |= ((a_addr >> 7) & 7u) << 17;  // base_offset field (bits 49-51 of the descriptor)
            const uint32_t alo = ahi + (Cfg::A_BYTES >> 4);
            const uint32_t bhi = kSmemDescLoConst | (((sS + offB) & 0x3FFFFu) >> 4);
            const uint32_t blo = bhi + (Cfg::B_BYTES >> 4);
            const int kbs = kb - kb0;  // 0: first k-block of this accumulator
            if (L.wide) {
#pragma unroll
              for (int k = 0; k < CBK / 16; ++k) {
                umma_f16(d_tmem, desc64(HIA, ahi + 2 * k), desc64(HI, bhi + 2 * k), IDESC2, (kbs | k) != 0 ? 1u : 0u);
                umma_f16(d_tmem, desc64(HIA, alo + 2 * k), desc64(HI, bhi + 2 * k), IDESC, 1u);
              }
            } else if (L.x3) {
#pragma unroll
              for (int k = 0; k < CBK / 16; ++k) {
                umma_f16(d_tmem, desc64(HIA, ahi + 2 * k), desc64(HI, bhi + 2 * k), IDESC, (kbs | k) != 0 ? 1u : 0u);
                umma_f16(d_tmem, desc64(HIA, ahi + 2 * k), desc64(HI, blo + 2 * k), IDESC, 1u);
                umma_f16(d_tmem, desc64(HIA, alo + 2 * k), desc64(HI, bhi + 2 * k), IDESC, 1u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < CBK / 16; ++k)
                umma_f16(d_tmem, desc64(HIA, ahi + 2 * k), desc64(HI, bhi + 2 * k), IDESC, (kbs | k) != 0 ? 1u : 0u);
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * slot);
        }
      }
    }
  } else {
    // 8 epilogue warps = 2 groups of 4 (one warp per TMEM lane quadrant); group g drains this CTA's tiles u with (u & 1) == g
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int rows_img = L.TW * L.TH;
    const int tb = row / rows_img;
    const int rr = row - tb * rows_img;
    const int ty = rr / L.TW, tx = rr - ty * L.TW;
    const bool row_ok = tb < L.TB;
    constexpr int NV = (NT % 32 == 0) ? 32 : 16;
    constexpr int NCHK = NT / NV;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t u = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if constexpr (!SEG) {
        if ((int)(u & 1u) != grp) { ++u; continue; }
      }
      const int c = (int)(t / L.tiles_per_cls);
      long long r = t - (long long)c * L.tiles_per_cls;
      const int nt = (int)(r % L.n_tiles_n);
      r /= L.n_tiles_n;
      const int ti = (int)(r % tiles_img);
      const int bt = (int)(r / tiles_img);
      const int sx = (ti % L.tiles_x) * L.TW + tx, sy = (ti / L.tiles_x) * L.TH + ty;
      const long long b = (long long)bt * L.TB + tb;
      const TcClass cl = L.cls[c];
      const bool ok = row_ok && b < L.B && sx < L.SW && sy < L.SH;
      int oy = cl.oy0 + cl.osy * sy, ox = cl.ox0 + cl.osx * sx;
      int cbase = nt * NT, boff = 0;
      if (L.nt_pixel_mode) {  // Dense -> Reshape(4,4,C): nt_pixel_mode N tiles per output pixel, bias indexed by the flat (pixel, channel)
        const int pix = nt / L.nt_pixel_mode;
        oy = pix / L.o.OW;
        ox = pix - oy * L.o.OW;
        cbase = (nt - pix * L.nt_pixel_mode) * NT;
        boff = pix * L.nt_pixel_mode * NT;
      }
      if constexpr (SEG) {
        // this group's chunks: q = grp, grp + 2, ...; partial sums promoted segment by segment
        constexpr int NMINE = (NCHK + 1) / 2;
        float acc[NMINE][NV];
#pragma unroll
        for (int i = 0; i < NMINE; ++i)
#pragma unroll
          for (int j = 0; j < NV; ++j) acc[i][j] = 0.f;
        const int nkb = cl.nkb;
        for (int kb0 = 0; kb0 < nkb; kb0 += L.seg_kb, ++u) {
          const uint32_t slot = u & (nslot - 1);
          const uint32_t tcol = lane_base + slot * slot_pitch;
          mbar_wait(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
          tc_fence_after();
#pragma unroll
          for (int i = 0; i < NMINE; ++i) {
            const int q = grp + 2 * i;
            if (q < NCHK) {
              float v[NV];
              if (L.wide) {
                float w[NV];
                tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
                tmem_ld_issue<NV>(tcol + (uint32_t)(NT + q * NV), w);
                tmem_ld_wait<NV>(v);
                tmem_ld_wait<NV>(w);
#pragma unroll
                for (int j = 0; j < NV; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fadd_rn(v[j], w[j]));
              } else {
                tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
                tmem_ld_wait<NV>(v);
#pragma unroll
                for (int j = 0; j < NV; ++j) acc[i][j] = __fadd_rn(acc[i][j], v[j]);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
        }
#pragma unroll
        for (int i = 0; i < NMINE; ++i) {
          const int q = grp + 2 * i;
          if (q < NCHK && ok) {
            ActRegs<NV> ra;
            act_prefetch<NV>(L.o, ok, oy, ox, cbase + q * NV, boff, ra);
            act_apply<NV>(L.o, oy, ox, cbase + q * NV, boff, ra, acc[i]);
            store_act<NV>(L.o, b, oy, ox, cbase + q * NV, acc[i]);
          }
        }
      } else {
      const uint32_t slot = u & (nslot - 1);
      const uint32_t tcol = lane_base + slot * slot_pitch;
      ActRegs<NV> ra;
      act_prefetch<NV>(L.o, ok, oy, ox, cbase, boff, ra);
      mbar_wait(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < NCHK; ++q) {
        if (q) act_prefetch<NV>(L.o, ok, oy, ox, cbase + q * NV, boff, ra);
        float v[NV];
        if (L.wide) {  // + the A_hi x B_lo partial product held in the second half of the tile's columns
          float w[NV];
          tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
          tmem_ld_issue<NV>(tcol + (uint32_t)(NT + q * NV), w);
          tmem_ld_wait<NV>(v);
          tmem_ld_wait<NV>(w);
#pragma unroll
          for (int j = 0; j < NV; ++j) v[j] += w[j];
        } else {
          tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
          tmem_ld_wait<NV>(v);
        }
        if (ok) {
          act_apply<NV>(L.o, oy, ox, cbase + q * NV, boff, ra, v);
          store_act<NV>(L.o, b, oy, ox, cbase + q * NV, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
      ++u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr int TC_MAX_SMEM = 232448;  // 227 KB

template <int CBK, int NT, bool SEG = false>
static int launch_one(const TcLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_conv_kernel<CBK, NT, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM);
  });
  if (attr_err != cudaSuccess)
    return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_conv_kernel<%d,%d>): %s", CBK, NT, cudaGetErrorString(attr_err));
  long long grid = L.total_tiles < max_ctas ? L.total_tiles : max_ctas;
  if (grid <= 0) return DBV_OK;
  const int smem = L.stages * L.stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
  if (L.stages < 2 || L.stages > 8 || smem > TC_MAX_SMEM) return fail(DBV_ERR_STATE, "tc_conv_kernel<%d,%d>: bad stage plan (%d x %d B)", CBK, NT, L.stages, L.stage_bytes);
  launch_pdl(tc_conv_kernel<CBK, NT, SEG>, (unsigned)grid, TC_THREADS, smem, st, L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

// fills stages / stage_bytes / wide from x3, CBK, NT
void tc_stage_plan(TcLayer& L, int CBK, int NT) {
  const int parts = L.x3 ? 2 : 1;
  L.stage_bytes = parts * (128 * CBK * 2 + NT * CBK * 2);
  int st = (TC_MAX_SMEM - 1536) / L.stage_bytes;
  L.stages = st > 8 ? 8 : st;
  L.wide = (L.x3 && 2 * NT <= 256) ? 1 : 0;
}

bool tc_layer_supported(int CBK, int NT) {
  if (CBK == 32) return NT == 16 || NT == 32 || NT == 64 || NT == 256;
  if (CBK == 64) return NT == 32 || NT == 64 || NT == 112 || NT == 128 || NT == 256;
  return false;
}

bool tc_seg_supported(int CBK, int NT) { return CBK == 64 && (NT == 64 || NT == 112 || NT == 128); }

int launch_tc_layer(const TcLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
  if (L.seg_kb > 0) {  // promoted partial sums (DBV_PREC_FP32TC)
    if (!L.wide) return fail(DBV_ERR_STATE, "segmented accumulation runs the hi/lo split layout with 2*NT <= 256 columns");
    if (CBK == 64 && NT == 64) return launch_one<64, 64, true>(L, max_ctas, st);
    if (CBK == 64 && NT == 112) return launch_one<64, 112, true>(L, max_ctas, st);
    if (CBK == 64 && NT == 128) return launch_one<64, 128, true>(L, max_ctas, st);
    return fail(DBV_ERR_UNSUPPORTED, "no segmented tcgen05 kernel instance for CBK=%d NT=%d", CBK, NT);
  }
#define DBV_TC_CASE(cb, nt) \
  if (CBK == cb && NT == nt) return launch_one<cb, nt>(L, max_ctas, st);
  DBV_TC_CASE(32, 16)
  DBV_TC_CASE(32, 32)
  DBV_TC_CASE(32, 64)
  DBV_TC_CASE(32, 256)
  DBV_TC_CASE(64, 32)
  DBV_TC_CASE(64, 64)
  DBV_TC_CASE(64, 112)
  DBV_TC_CASE(64, 128)
  DBV_TC_CASE(64, 256)
#undef DBV_TC_CASE
  return fail(DBV_ERR_UNSUPPORTED, "no tcgen05 kernel instance for CBK=%d NT=%d", CBK, NT);
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes, int elem_bytes) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      return fail(DBV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(DBV_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u] swz=%d",
                (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
                (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
                (unsigned long long)(rank > 4 ? gdims[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, swizzle_bytes);
  }
  return DBV_OK;
}

}  // namespace dbv
