// tcgen05 implicit-GEMM convolution with a RESIDENT HALO TILE (sm_100a).
//
// A streaming kernel re-reads the activation tile from L2 once per 3x3 tap (and, in the hi/lo split precisions, once
// per plane), which made the large-image / few-channel layers L2-bandwidth bound (profiles/r01_v1_summary.md).  Here a
// CTA loads a band of R+2 input rows ONCE per plane (TMA, out-of-bounds zero fill = TF "SAME" padding) and all taps,
// all output-parity classes of a stride-2 transposed conv, all parity planes of a stride-2 conv and all hi/lo pairings
// read it as *row-shifted windows of the same shared-memory tile*: with row pitch WP = W + 1 (slot 0 of a row is x = -1,
// shared as the right neighbour of the previous row's last pixel) output position p = y*WP + x of the band reads, for
// tap (dy, dx), shared-memory row p + (dy+1)*WP + (dx+1).  That works because tcgen05 applies the 64B/128B swizzle to
// absolute shared-memory address bits, so a descriptor whose start is shifted by whole rows stays consistent with what
// TMA wrote (tests/test_gpu_network.py::test_probe_descriptor_row_shift).  Weights are loaded once per CTA and stay
// resident.  One position per row (x = W) computes garbage that the epilogue drops.
//
//   warp 0    TMA producer (weights once; one halo band per work item, ring of nbuf buffers)
//   warp 1    MMA issuer: per unit (several (class, tile) sub-units sharing one TMEM slot) a flat, host-made list of MMAs
//   warps 2-9 epilogue, two groups of 4: hi/lo accumulator sum, bias, PReLU(h,w,c) / ReLU, 16-bit hi[/lo] conversion, stores
#include "tc_ptx.cuh"
#include <mutex>
#include <type_traits>

#ifdef DBV_ABLATE
#include "../../include/debvader_b200_debug.h"
// clock64 instrumentation (ablation build only): cycles per role, summed over CTAs / warps with atomics at kernel end
//  0 mma: total            1 mma: waiting for a halo band (afull)     2 mma: waiting for a free accumulator slot (tempty)
//  3 mma: issuing (list walk + tcgen05.mma + commits)                 4 mma: MMAs issued        5 mma: units
//  6 tma: total            7 tma: waiting for a free halo buffer (aempty)
//  8 epi: total (sum over the epilogue warps)   9 epi: waiting for the accumulators (tfull)   10 epi: tcgen05.ld issue -> wait::ld return
// 11 epi: alpha load issue (address math + ld)  12 epi: math + stores   13 epi: items   14 CTAs   15 epi: slot release (fence + arrive)
__device__ unsigned long long g_halo_ctr[DBV_HALO_NLAYERS][DBV_HALO_NCOUNTERS];  // [layer index (kLayers)][counter]
#define HCLK() clock64()
#define HADD(i, v) atomicAdd(&g_halo_ctr[L.dbg_id][i], (unsigned long long)(v))
#else
#define HCLK() 0ll
#define HADD(i, v) ((void)0)
#endif

namespace dbv {

#ifndef DBV_HALO_EPI_GROUPS
#define DBV_HALO_EPI_GROUPS 2
#endif
constexpr int HALO_EPI_GROUPS = DBV_HALO_EPI_GROUPS;  // epilogue groups of 4 warps (one warp per TMEM lane quadrant); 2 x 32-channel items measured best
constexpr int HALO_THREADS = 64 + HALO_EPI_GROUPS * 128;  // TMA warp, MMA warp, epilogue groups
constexpr int HALO_NBUF_MAX = 4;           // halo buffers in the ring (barrier storage)
constexpr int HALO_NSLOT_MAX = 8;          // accumulator slots in the TMEM ring (512 columns / slot width, capped)

// NOSWZ (encoder conv1, Cin = 6 padded to 8): one 16-byte row per pixel, no swizzle.  A K=16 MMA operand is
// then TWO ADJACENT PIXELS: core-matrix stride along K (LBO) = 16 bytes = the pixel pitch, so the 3x3x8
// im2col never exists anywhere — the (kx, channel) axis of each kernel row is read as overlapping
// windows of the halo tile.  Per kernel row ky: pixels (x-1, x) and (x+1, x+2[zero weights]).
//
// Work decomposition: item g in [0, B * bands_per_img) is band  g / B  of stamp  g % B  (band-major), and CTA i
// owns the contiguous range [total*i/grid, total*(i+1)/grid): a CTA stays on ONE band row for (almost) its whole
// life, so the PReLU alpha slice of that band (tens of KB, the same for every stamp) stays L1-resident.
//
// Accumulators: TMEM is a ring of 512 / (U*DW) slots; a "unit" = up to U consecutive sub-units (one 128-position tile of
// one output-parity class each) of a band.  The MMA warp issues all MMAs of a unit into the next free slot and commits
// it once; the epilogue groups split the unit's (sub-unit, channel chunk) items and all release the slot.
template <int CBK, int NT, bool NOSWZ, bool SEGS>
__global__ void __launch_bounds__(HALO_THREADS, 1) tc_halo_kernel(const __grid_constant__ HaloLayer L) {
  pdl_trigger();
  constexpr bool CG8 = !NOSWZ && CBK == 16;             // channel-group-planar input: 16-byte pixel rows, K=16 = two groups one region apart
  constexpr int ROWB = (NOSWZ || CG8) ? 16 : CBK * 2;   // activation row pitch in shared memory
  constexpr int ROWB_W = NOSWZ ? 16 : CBK * 2;          // weight rows (K-major, swizzled; conv1: host-packed core matrices)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                              // nbuf x n_regions x region_bytes
  const uint32_t sW = sA + L.nbuf * L.buf_bytes;  // resident weights: n_wblk blocks of NT x ROWB
  const uint32_t sBar = sW + L.w_bytes + L.tail_pad;
  const uint32_t bar_w = sBar, bar_afull = sBar + 8, bar_aempty = sBar + 40, bar_tfull = sBar + 72, bar_tempty = sBar + 136;
  const uint32_t s_tmem = sBar + 200;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A "unit" = the sub-units (one 128-position tile of one output class — or of all four classes of a stride-2 transposed
  // conv, concatenated along N) that share ONE accumulator slot, one tempty wait and one commit.  The host lays the whole
  // band out as flat tables: ops[] (one entry per tcgen05.mma, cut into units by unit_op_end[]) and items[] (one entry per
  // epilogue item, cut by unit_item_end[]).
  const uint32_t SW = (uint32_t)L.slot_cols;           // slot width (power of two)
  const uint32_t nslot = (512u / SW) < (uint32_t)HALO_NSLOT_MAX ? (512u / SW) : (uint32_t)HALO_NSLOT_MAX;  // power of two
  const uint32_t slot_shift = 31u - (uint32_t)__clz((int)nslot);
  const int units_per_band = L.n_units;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    if (!NOSWZ) tma_prefetch_desc(&L.tmB);
    mbar_init(bar_w, 1);
    for (int s = 0; s < HALO_NBUF_MAX; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < HALO_NSLOT_MAX; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4 * HALO_EPI_GROUPS);  // every epilogue warp releases every unit
    }
    fence_barrier_init();
  }
  // zero the slack after every region's TMA box once: with the W+1 pitch the slot after the last halo row is the right
  // neighbour of its last pixel and must read as zero padding (TMA never writes there)
  for (int r = 0; r < L.nbuf * L.n_regions; ++r)
    for (int i = L.a_box_bytes + 16 * (int)threadIdx.x; i < L.region_bytes; i += 16 * HALO_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + (r / L.n_regions) * L.buf_bytes + (r % L.n_regions) * L.region_bytes + i), "r"(0) : "memory");
  for (int r = 0; r < L.nbuf; ++r)  // ... and the slack after each buffer (CG8: the regions themselves are packed)
    for (int i = L.n_regions * L.region_bytes + 16 * (int)threadIdx.x; i < L.buf_bytes; i += 16 * HALO_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + r * L.buf_bytes + i), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // nothing produced or consumed by the previous kernel is touched above this line
  const long long total = L.total_bands;
  const long long g0 = total * blockIdx.x / gridDim.x, g1 = total * (blockIdx.x + 1) / gridDim.x;
  const long long yb0 = g0 / L.B;
  const int b_first = (int)(g0 - yb0 * L.B);

  if (warp == 0) {
    if (elect_one()) {
      if constexpr (NOSWZ) {
        mbar_expect_tx(bar_w, (uint32_t)L.w_bytes);
        bulk_load(sW, L.w_img, (uint32_t)L.w_bytes, bar_w);
      } else {
        mbar_expect_tx(bar_w, (uint32_t)(L.n_wblk * NT * ROWB_W));
        for (int blk = 0; blk < L.n_wblk; ++blk) tma_load_2d(sW + blk * (NT * ROWB_W), &L.tmB, bar_w, 0, (int)L.w_src[blk] * L.w_rows_per_blk);
      }
      int stage = 0;
      uint32_t phase = 0;
      int b = b_first, y0 = (int)yb0 * L.R;
      [[maybe_unused]] const long long tk0 = HCLK();
      [[maybe_unused]] long long tk_wait = 0;
      for (long long g = g0; g < g1; ++g) {
        [[maybe_unused]] const long long tka = HCLK();
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        tk_wait += HCLK() - tka;
        mbar_expect_tx(bar_afull + 8 * stage, (uint32_t)(L.n_regions * L.a_box_bytes));
        if constexpr (CG8) {  // one box: all channel-group planes, whole rows of 16-byte pixels (x in u64 units)
          tma_load_5d(sA + stage * L.buf_bytes, &L.tmA, bar_afull + 8 * stage, -2 * L.pad, y0 - L.pad_top, 0, b, 0);
        } else {
          for (int r = 0; r < L.n_regions; ++r)
            tma_load_5d(sA + stage * L.buf_bytes + r * L.region_bytes, &L.tmA, bar_afull + 8 * stage, L.region_coff[r], -L.pad, y0 - L.pad_top, L.region_c3[r], b);
        }
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
        if (++b == (int)L.B) { b = 0; y0 += L.R; }
      }
      HADD(6, HCLK() - tk0);
      HADD(7, tk_wait);
      HADD(14, 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(bar_w, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;  // running unit counter (slot = u % nslot)
      // descriptor words: swizzled K-major rows, or (NOSWZ) interleaved 8x16-byte core matrices with
      // A: SBO 128 B (8 pixels), LBO 16 B (next pixel);  B: SBO 256 B, LBO 128 B (host-packed image)
      constexpr uint32_t HI = (NOSWZ || CG8) ? ((128u >> 4) | (1u << 14)) : smem_desc_hi<ROWB>();
      constexpr uint32_t HIB = NOSWZ ? ((256u >> 4) | (1u << 14)) : smem_desc_hi<ROWB_W>();
      constexpr uint32_t LOB = NOSWZ ? ((128u >> 4) << 16) : kSmemDescLoConst;
      // A low word: LBO = distance (>>4) between the two K-halves of a K=16 operand: 16 B = the next pixel (conv1), one region =
      // the next channel group (CG8); unused (1) for swizzled rows
      const uint32_t LOA = CG8 ? ((uint32_t)(L.region_bytes >> 4) << 16) : kSmemDescLoConst;
      const uint32_t w16 = LOB | (sW >> 4);
      const int nunits = (DBV_DBG(L.dbg_skip) & 1) ? 0 : units_per_band;
      [[maybe_unused]] const long long mk0 = HCLK();
      [[maybe_unused]] long long mk_afull = 0, mk_tempty = 0, mk_n = 0;
      for (long long g = g0; g < g1; ++g) {
        [[maybe_unused]] const long long mka = HCLK();
        mbar_wait(bar_afull + 8 * stage, phase);
        tc_fence_after();
        mk_afull += HCLK() - mka;
        const uint32_t a16 = LOA | ((sA + stage * L.buf_bytes) >> 4);
        int i = 0;
        for (int k = 0; k < nunits; ++k, ++u) {
          const uint32_t slot = u & (nslot - 1);
          [[maybe_unused]] const long long mkb = HCLK();
          mbar_wait(bar_tempty + 8 * slot, ((u >> slot_shift) & 1u) ^ 1u);
          tc_fence_after();
          mk_tempty += HCLK() - mkb;
          const uint32_t d0 = tmem_base + slot * SW;
          const int iend = L.unit_op_end[k];
#ifdef DBV_ABLATE
          mk_n += iend - i;
#endif
          // flat list, one 16-byte entry per MMA read with a uniform constant load: the loop is pure issue
#pragma unroll 4
          for (; i < iend; ++i) {
            const HaloOp e = L.ops[i];
            umma_f16(d0 + (e.d & 0xffffu), desc64(HI, a16 + e.a), desc64(HIB, w16 + e.b), e.idesc, e.d >> 16);
          }
          umma_commit(bar_tfull + 8 * slot);
        }
        umma_commit(bar_aempty + 8 * stage);
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
      }
      {
        [[maybe_unused]] const long long tot = HCLK() - mk0;
        HADD(0, tot);
        HADD(1, mk_afull);
        HADD(2, mk_tempty);
        HADD(3, tot - mk_afull - mk_tempty);
        HADD(4, mk_n);
        HADD(5, u);
      }
    }
  } else {
    // Epilogue: HALO_EPI_GROUPS groups of 4 warps (one warp per TMEM lane quadrant).  Every group walks every unit and
    // takes every G-th item (sub-unit x NV-channel chunk) of it; all 4*G warps release the slot.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    constexpr int NV = (HALO_EPI_GROUPS > 3 || NT % 32 != 0) ? 16 : 32;  // channels per item: 16 keeps 16 epilogue warps spill-free
    static_assert(NT % NV == 0, "items cover whole channel chunks");
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int ncls = (DBV_DBG(L.dbg_skip) & 1) ? 0 : L.n_cls;  // units exist only if the MMA warp produces them
    const bool has_alpha = L.o.alpha != nullptr && !(DBV_DBG(L.dbg_skip) & 4);  // halo layers: PReLU(h,w,c) or (head) ReLU, never a second PReLU
    const float4* alpha4 = reinterpret_cast<const float4*>(L.o.alpha);
    const bool le1 = L.o.alpha_le1 != 0;
    const uint32_t npix = (uint32_t)(L.o.OH * L.o.OW);
    // The output layout (mode / planes / 16-bit format) is constant for a launch: the loops are instantiated once per
    // layout actually used, with those OutSpec fields as compile-time constants, so that store_act's dispatch folds away
    // (the epilogue is instruction-issue bound: ~250 instructions per 32-channel item).
    auto run = [&](auto MODE, auto PLANES, auto F16) {
      OutSpec o = L.o;
      if constexpr (decltype(MODE)::value >= 0) {
        o.mode = decltype(MODE)::value;
        o.planes = decltype(PLANES)::value;
        o.f16 = decltype(F16)::value;
      }
      uint32_t u = 0;
      int b = b_first, y0 = (int)yb0 * L.R;
      [[maybe_unused]] const long long ek0 = HCLK();
      [[maybe_unused]] long long ek_tfull = 0, ek_ld = 0, ek_alpha = 0, ek_math = 0, ek_items = 0, ek_rel = 0;
      for (long long g = g0; g < g1; ++g) {
        for (int k = 0; k < (ncls ? units_per_band : 0); ++k, ++u) {
          const uint32_t slot = u & (nslot - 1);
          bool waited = false;
          // items of this unit (host-made table: accumulator column, class, tile, channel chunk); group g takes items g, g + G, ...
          const int it0 = k ? (int)L.unit_item_end[k - 1] : 0, it1 = (int)L.unit_item_end[k];
  #pragma unroll 1
          for (int item = it0 + grp; item < it1; item += HALO_EPI_GROUPS) {
            const uint32_t it = L.items[item];
            const int c = (int)((it >> 9) & 3u), m = (int)((it >> 11) & 31u), q = (int)((it >> 16) & 7u);
            const int p = 128 * m + row;
            const int ly = (int)__umulhi((uint32_t)p, L.magic_wp), sx = p - ly * L.WP, sy = y0 + ly;
            const bool ok = ly < L.R && sx < L.W && sy < L.H && !(DBV_DBG(L.dbg_skip) & 2);
            const int oy = L.cls[c].oy0 + L.cls[c].osy * sy, ox = L.cls[c].ox0 + L.cls[c].osx * sx;
            const uint32_t tcol = lane_base + slot * SW + (it & 511u);  // first column of this item's NV channels
            const int c0 = q * NV;
            // PReLU slopes of this thread's pixel ([C/4][pixels][4] layout: 32-bit element offsets, one 16-byte load per 4 channels),
            // requested before the accumulator wait
            float4 al[NV / 4];
            [[maybe_unused]] const long long eka = HCLK();
            if (has_alpha && ok) {
              const uint32_t off = (uint32_t)(c0 >> 2) * npix + (uint32_t)(oy * L.o.OW + ox);
  #pragma unroll
              for (int j = 0; j < NV / 4; ++j) al[j] = __ldg(alpha4 + off + (uint32_t)j * npix);
            }
            [[maybe_unused]] const long long ekb = HCLK();
            if (!waited) {
              mbar_wait(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
              tc_fence_after();
              waited = true;
            }
            [[maybe_unused]] const long long ekc = HCLK();
            float v[NV];
            if (L.wide) {  // + the A_hi x B_lo partial product held in the second half of the sub-unit's columns
              float w[NV];
              tmem_ld_issue<NV>(tcol, v);
              tmem_ld_issue<NV>(tcol + (uint32_t)NT, w);
              tmem_ld_wait<NV>(v);
              tmem_ld_wait<NV>(w);
  #pragma unroll
              for (int j = 0; j < NV; j += 2) add2(v[j], v[j + 1], w[j], w[j + 1]);
              if constexpr (SEGS) {  // DBV_PREC_FP32TC: the other partial accumulators of this tile, promoted in fp32 registers
                for (int sg = 1; sg < L.nseg; ++sg) {
                  float w2[NV];
                  tmem_ld_issue<NV>(tcol + (uint32_t)(sg * L.seg_cols), w);
                  tmem_ld_issue<NV>(tcol + (uint32_t)(sg * L.seg_cols + NT), w2);
                  tmem_ld_wait<NV>(w);
                  tmem_ld_wait<NV>(w2);
  #pragma unroll
                  for (int j = 0; j < NV; ++j) v[j] = __fadd_rn(v[j], __fadd_rn(w[j], w2[j]));
                }
              }
            } else {
              tmem_ld_issue<NV>(tcol, v);
              tmem_ld_wait<NV>(v);
            }
            [[maybe_unused]] const long long ekd = HCLK();
            if (ok) {
              if (!(DBV_DBG(L.dbg_skip) & 4)) {
  #pragma unroll
                for (int j = 0; j < NV; ++j) v[j] += L.bias_c[c0 + j];  // bias from the constant bank
                if (has_alpha) {
  #pragma unroll
                  for (int j = 0; j < NV / 4; ++j) prelu4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], al[j], le1);
                } else if (L.o.relu) {
  #pragma unroll
                  for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
                }
              }
              if (!(DBV_DBG(L.dbg_skip) & 8)) store_act<NV>(o, b, oy, ox, c0, v);
              else if (v[0] == 123.456f) store_act<NV>(o, b, oy, ox, c0, v);  // keep the loads / math alive
            }
#ifdef DBV_ABLATE
            {
              const long long eke = HCLK();
              ek_alpha += ekb - eka;
              ek_tfull += ekc - ekb;
              ek_ld += ekd - ekc;
              ek_math += eke - ekd;
              ek_items += 1;
            }
#endif
          }
          if (!waited) {  // a group without a sub-unit in this unit still takes part in the hand-over, in phase order
            mbar_wait(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
            tc_fence_after();
          }
          [[maybe_unused]] const long long ekr = HCLK();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
          ek_rel += HCLK() - ekr;
        }
        if (++b == (int)L.B) { b = 0; y0 += L.R; }
      }
      if (lane == 0) {
        HADD(8, HCLK() - ek0);
        HADD(9, ek_tfull);
        HADD(10, ek_ld);
        HADD(11, ek_alpha);
        HADD(12, ek_math);
        HADD(13, ek_items);
        HADD(15, ek_rel);
      }
    };
    using std::integral_constant;
    const int om = L.o.mode, op = L.o.planes, of = L.o.f16;
    if (om == OUT_BF16_NHWC && op == 2 && of == 0) run(integral_constant<int, OUT_BF16_NHWC>{}, integral_constant<int, 2>{}, integral_constant<int, 0>{});
    else if (om == OUT_BF16_NHWC && op == 1 && of == 1) run(integral_constant<int, OUT_BF16_NHWC>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{});
    else if (om == OUT_BF16_PARITY && op == 2 && of == 0) run(integral_constant<int, OUT_BF16_PARITY>{}, integral_constant<int, 2>{}, integral_constant<int, 0>{});
    else if (om == OUT_HEAD) run(integral_constant<int, OUT_HEAD>{}, integral_constant<int, 1>{}, integral_constant<int, 0>{});
    else run(integral_constant<int, -1>{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});  // any other layout: runtime dispatch
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int CBK, int NT, bool NOSWZ = false, bool SEGS = false>
static int launch_halo_one(const HaloLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_halo_kernel<CBK, NT, NOSWZ, SEGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_MAX_SMEM);
  });
  if (attr_err != cudaSuccess)
    return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_halo_kernel<%d,%d>): %s", CBK, NT, cudaGetErrorString(attr_err));
  const long long grid = L.total_bands < max_ctas ? L.total_bands : max_ctas;
  if (grid <= 0) return DBV_OK;
  launch_pdl(tc_halo_kernel<CBK, NT, NOSWZ, SEGS>, (unsigned)grid, HALO_THREADS, L.smem_bytes, st, L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

}  // namespace dbv
#ifdef DBV_ABLATE
extern "C" int dbv_halo_counters(unsigned long long* out_host, int reset) {
  if (out_host) {
    cudaError_t e = cudaMemcpyFromSymbol(out_host, g_halo_ctr, sizeof(g_halo_ctr));
    if (e != cudaSuccess) return dbv::fail(DBV_ERR_CUDA, "dbv_halo_counters: %s", cudaGetErrorString(e));
  }
  if (reset) {
    static unsigned long long z[DBV_HALO_NLAYERS][DBV_HALO_NCOUNTERS] = {};
    cudaError_t e = cudaMemcpyToSymbol(g_halo_ctr, z, sizeof(z));
    if (e != cudaSuccess) return dbv::fail(DBV_ERR_CUDA, "dbv_halo_counters: %s", cudaGetErrorString(e));
  }
  return DBV_OK;
}
#endif
namespace dbv {

bool halo_layer_supported(int CBK, int NT) {
  if (CBK == 16) return NT == 16 || NT == 32 || NT == 64;  // conv1 (no-swizzle pair trick, NT = 32) / channel-group-planar inputs
  if (CBK == 32) return NT == 16 || NT == 32 || NT == 64;
  if (CBK == 64) return NT == 32 || NT == 64 || NT == 128;
  return false;
}

int launch_halo_layer(const HaloLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
  if (L.nseg > 1) {  // several partial accumulators per tile (DBV_PREC_FP32TC): the instances whose epilogue adds them up
    if (L.cg8 || !L.wide) return fail(DBV_ERR_STATE, "segmented halo accumulators: hi/lo split layers with NHWC inputs only");
    if (CBK == 32 && NT == 16) return launch_halo_one<32, 16, false, true>(L, max_ctas, st);
    if (CBK == 32 && NT == 32) return launch_halo_one<32, 32, false, true>(L, max_ctas, st);
    if (CBK == 32 && NT == 64) return launch_halo_one<32, 64, false, true>(L, max_ctas, st);
    if (CBK == 64 && NT == 64) return launch_halo_one<64, 64, false, true>(L, max_ctas, st);
    return fail(DBV_ERR_UNSUPPORTED, "no segmented halo kernel instance for CBK=%d NT=%d", CBK, NT);
  }
#define DBV_HALO_CASE(cb, nt) \
  if (CBK == cb && NT == nt) return launch_halo_one<cb, nt>(L, max_ctas, st);
  if (CBK == 16 && L.cg8) {
    if (NT == 16) return launch_halo_one<16, 16>(L, max_ctas, st);
    if (NT == 32) return launch_halo_one<16, 32>(L, max_ctas, st);
    if (NT == 64) return launch_halo_one<16, 64>(L, max_ctas, st);
  }
  if (CBK == 16 && NT == 32) return launch_halo_one<16, 32, true>(L, max_ctas, st);
  DBV_HALO_CASE(32, 16)
  DBV_HALO_CASE(32, 32)
  DBV_HALO_CASE(32, 64)
  DBV_HALO_CASE(64, 32)
  DBV_HALO_CASE(64, 64)
  DBV_HALO_CASE(64, 128)
#undef DBV_HALO_CASE
  return fail(DBV_ERR_UNSUPPORTED, "no halo kernel instance for CBK=%d NT=%d", CBK, NT);
}

}  // namespace dbv
