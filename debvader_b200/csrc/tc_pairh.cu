// CTA-pair tcgen05 convolution with a PER-CHUNK HALO BOX (sm_100a) for the 8x8 .. 16x16, 128/256-channel layers
// (enc_conv5, enc_conv7, dec_convT2, dec_convT3, dec_convT4).
//
// tc_pair.cu re-loads the 128-row activation tile once per 3x3 tap, so these layers stream 9 x (A_hi + A_lo) + the
// weight halves per (tap, channel chunk) and are bound by L2 -> shared-memory fill (measured ~7 TB/s aggregate), not by
// the tensor pipe.  Here the activation operand of one 64-channel chunk is loaded ONCE as a halo box and every tap reads
// it as a shifted window:
//
//   box (64 ch, 10 px, TB stamps, HB rows), tensor-map dimension order (C, W, B, H)  ->  shared memory [row y][stamp][x][c]
//
// An MMA operand row group (8 rows = one 128B-swizzle atom) is then the 8 pixels x0..x0+7 of one image row of one stamp,
// consecutive groups (y, stamp) are exactly 10 pixel rows = 1280 bytes apart — a uniform stride, which is all the
// K-major descriptor needs (SBO = 1280) — and tap (dy, dx) is the same window moved by (dy * TB * 10 + dx) pixel rows.
// (tcgen05 applies the swizzle to absolute shared-memory address bits, so windows that start on any 128-byte row stay
// consistent with what TMA wrote: tests/test_gpu_network.py::test_probe_descriptor_row_shift.)
// Tiles are 8 pixels wide: 8x8 maps pair two stamps per CTA (TB = 2), 15/16-row maps use one 8 x 16 strip (TB = 1).
//
// Pipeline per cluster work item (a pair of 128-row tiles, one per CTA; cta_group::2, M = 256):
//   for chunk:  A ring  <- halo box hi, lo of the CTA's own tile                        (both CTAs, leader's barrier)
//     for class, tap:  B ring <- this CTA's half (NT/2 rows) of the weight block hi, lo
//        D[class] += A_hi(tap) B_hi + A_hi(tap) B_lo + A_lo(tap) B_hi                  (leader, 3 x 4 MMAs of K = 16)
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"
#include <mutex>

namespace dbv {

constexpr int PH_THREADS = 64 + 2 * 128;
constexpr int PH_MAX_SMEM = 232448;

template <int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PH_THREADS, 1) tc_pairh_kernel(const __grid_constant__ PairHLayer L) {
  pdl_trigger();
  constexpr int ROWB = 128;                  // 64 channels
  constexpr int BH_BYTES = (NT / 2) * ROWB;  // this CTA's half of one weight block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_stage = 2u * (uint32_t)L.abox_bytes;  // hi, lo boxes (each rounded up to 1024)
  const uint32_t b_stage = 2u * BH_BYTES;
  const uint32_t sA = base, sB = base + (uint32_t)L.a_stages * a_stage;
  const uint32_t sBar = sB + (uint32_t)L.b_stages * b_stage + (uint32_t)L.tail_pad;
  const uint32_t bar_afull = sBar, bar_aempty = sBar + 32, bar_bfull = sBar + 64, bar_bempty = sBar + 128;
  const uint32_t bar_tfull = sBar + 192, bar_tempty = sBar + 208, s_tmem = sBar + 224;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int gs = L.n_cls / L.cls_groups;                                   // classes per work item
  const uint32_t acc_cols = (uint32_t)(gs * NT);                           // accumulator columns of one item
  const uint32_t nslot = acc_cols <= 256 ? 2u : 1u;                        // accumulator double buffering when it fits

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    tma_prefetch_desc(&L.tmB);
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 16);  // 8 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // nothing produced or consumed by the previous kernel is touched above this line

  const long long items = L.pair_items * L.cls_groups;  // (tile pair, class group)
  const uint32_t cid = cluster_id_x(), ncl = nclusters_x();
  const int tiles_img = L.tiles_x;  // one strip row of 8-pixel-wide tiles per image group

  if (warp == 0) {
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t atx = 2u * 2u * (uint32_t)L.abox_tx, btx = 2u * 2u * (uint32_t)BH_BYTES;  // both CTAs, hi + lo
      const uint32_t afull0 = map_to_rank(bar_afull, 0), bfull0 = map_to_rank(bar_bfull, 0);
      for (long long w = cid; w < items; w += ncl) {
        const long long m = 2 * (w / L.cls_groups) + rank;  // this CTA's tile (may be one past the end: all-zero loads, masked stores)
        const int cg = (int)(w % L.cls_groups);
        const int t0 = L.cls_begin[cg * gs], t1 = L.cls_begin[(cg + 1) * gs];
        const int ti = (int)(m % tiles_img);
        const int b0 = (int)(m / tiles_img) * L.TB;
        const int x0 = ti * 8 - 1;
        for (int ch = 0; ch < L.nchunk; ++ch) {
          mbar_wait_cluster(bar_aempty + 8 * as, aph ^ 1u);
          if (rank == 0) mbar_expect_tx(bar_afull + 8 * as, atx);
          const uint32_t dA = sA + (uint32_t)as * a_stage;
          tma2_load_5d(dA, &L.tmA, afull0 + 8 * as, ch * 64, x0, b0, -1, 0);
          tma2_load_5d(dA + (uint32_t)L.abox_bytes, &L.tmA, afull0 + 8 * as, L.lo_coff + ch * 64, x0, b0, -1, 0);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
          for (int t = t0; t < t1; ++t) {
            mbar_wait_cluster(bar_bempty + 8 * bs, bph ^ 1u);
            if (rank == 0) mbar_expect_tx(bar_bfull + 8 * bs, btx);
            const uint32_t dB = sB + (uint32_t)bs * b_stage;
            const int brow = L.tap_brow[t] + ch * L.chunk_brow + (int)rank * (NT / 2);
            tma2_load_2d(dB, &L.tmB, bfull0 + 8 * bs, 0, brow);
            tma2_load_2d(dB + BH_BYTES, &L.tmB, bfull0 + 8 * bs, 0, brow + L.lo_brow);
            if (++bs == L.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, u = 0;
      const uint32_t IDESC = (1u << 4) | idesc_ab_fmt(L.ab_f16) | ((uint32_t)(NT >> 3) << 17) | ((256u >> 4) << 24);
      // A: 128B-swizzled K-major rows, 8-row groups 10 pixel rows (1280 B) apart;  B: dense groups (1024 B)
      constexpr uint32_t HIA = (uint32_t)(1280 >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t HIB = smem_desc_hi<ROWB>();
      for (long long w = cid; w < items; w += ncl, ++u) {
        const uint32_t slot = nslot == 2 ? (u & 1u) : 0u;
        const uint32_t tph = nslot == 2 ? ((u >> 1) & 1u) : (u & 1u);
        mbar_wait_cluster(bar_tempty + 8 * slot, tph ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + slot * 256u;
        const int c_first = (int)(w % L.cls_groups) * gs;
        for (int ch = 0; ch < L.nchunk; ++ch) {
          mbar_wait_cluster(bar_afull + 8 * as, aph);
          tc_fence_after();
          const uint32_t aBase = sA + (uint32_t)as * a_stage;
          for (int c = c_first; c < c_first + gs; ++c) {
            const uint32_t d = d0 + (uint32_t)(c - c_first) * NT;
            for (int t = L.cls_begin[c]; t < L.cls_begin[c + 1]; ++t) {
              mbar_wait_cluster(bar_bfull + 8 * bs, bph);
              tc_fence_after();
              const uint32_t ahi = kSmemDescLoConst | (((aBase + (uint32_t)L.tap_aoff[t]) & 0x3FFFFu) >> 4);
              const uint32_t alo = ahi + ((uint32_t)L.abox_bytes >> 4);
              const uint32_t bhi = kSmemDescLoConst | (((sB + (uint32_t)bs * b_stage) & 0x3FFFFu) >> 4);
              const uint32_t blo = bhi + (BH_BYTES >> 4);
              const uint32_t first = (ch == 0 && t == L.cls_begin[c]) ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma2_f16(d, desc64(HIA, ahi + 2 * k), desc64(HIB, bhi + 2 * k), IDESC, (k == 0) ? first : 1u);
                umma2_f16(d, desc64(HIA, ahi + 2 * k), desc64(HIB, blo + 2 * k), IDESC, 1u);
                umma2_f16(d, desc64(HIA, alo + 2 * k), desc64(HIB, bhi + 2 * k), IDESC, 1u);
              }
              umma2_commit_mc(bar_bempty + 8 * bs);
              if (++bs == L.b_stages) { bs = 0; bph ^= 1u; }
            }
          }
          umma2_commit_mc(bar_aempty + 8 * as);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
        }
        umma2_commit_mc(bar_tfull + 8 * slot);
      }
    }
  } else {
    // accumulator row r: 8-row group j = r / 8 is (image row ty = j / TB, stamp tb = j % TB), pixel tx = r % 8
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int j = row >> 3, tx = row & 7;
    const int tb = j % L.TB, ty = j / L.TB;
    constexpr int NV = 32;
    constexpr int NCHK = NT / NV;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tempty0 = map_to_rank(bar_tempty, 0);
    uint32_t u = 0;
    for (long long w = cid; w < items; w += ncl, ++u) {
      const long long m = 2 * (w / L.cls_groups) + rank;
      const int c_first = (int)(w % L.cls_groups) * gs;
      const int ti = (int)(m % tiles_img);
      const long long b = (m / tiles_img) * L.TB + tb;
      const int sx = ti * 8 + tx, sy = ty;
      const bool ok = b < L.B && sx < L.SW && sy < L.SH;
      const uint32_t slot = nslot == 2 ? (u & 1u) : 0u;
      const uint32_t tph = nslot == 2 ? ((u >> 1) & 1u) : (u & 1u);
      bool waited = false;
      for (int c = c_first; c < c_first + gs; ++c) {
        const int oy = L.cls[c].oy0 + L.cls[c].osy * sy, ox = L.cls[c].ox0 + L.cls[c].osx * sx;
        const uint32_t tcol = lane_base + slot * 256u + (uint32_t)(c - c_first) * NT;
#pragma unroll 1
        for (int q = grp; q < NCHK; q += 2) {  // the two groups split the 32-channel chunks
          ActRegs<NV> ra;
          act_prefetch<NV>(L.o, ok, oy, ox, q * NV, 0, ra);
          if (!waited) {
            mbar_wait_cluster_relaxed(bar_tfull + 8 * slot, tph);
            tc_fence_after();
            waited = true;
          }
          float v[NV];
          tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
          tmem_ld_wait<NV>(v);
          if (ok) {
            act_apply<NV>(L.o, oy, ox, q * NV, 0, ra, v);
            store_act<NV>(L.o, b, oy, ox, q * NV, v);
          }
        }
      }
      if (!waited) {
        mbar_wait_cluster_relaxed(bar_tfull + 8 * slot, tph);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + 8 * slot);  // nothing to publish through memory: see tc_pair_ptx.cuh
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int NT>
static int launch_pairh_one(const PairHLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_pairh_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, PH_MAX_SMEM); });
  if (attr_err != cudaSuccess) return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_pairh_kernel<%d>): %s", NT, cudaGetErrorString(attr_err));
  if (L.pair_items <= 0) return DBV_OK;
  const long long clusters = L.pair_items < max_ctas / 2 ? L.pair_items : max_ctas / 2;
  if (L.smem_bytes > PH_MAX_SMEM) return fail(DBV_ERR_STATE, "tc_pairh_kernel<%d>: %d bytes of shared memory", NT, L.smem_bytes);
  launch_pdl(tc_pairh_kernel<NT>, (unsigned)(2 * clusters), PH_THREADS, L.smem_bytes, st, L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

bool tc_pairh_supported(int NT) { return NT == 64 || NT == 128 || NT == 256; }

int launch_tc_pairh(const PairHLayer& L, int NT, int max_ctas, cudaStream_t st) {
  if (NT == 64) return launch_pairh_one<64>(L, max_ctas, st);
  if (NT == 128) return launch_pairh_one<128>(L, max_ctas, st);
  if (NT == 256) return launch_pairh_one<256>(L, max_ctas, st);
  return fail(DBV_ERR_UNSUPPORTED, "no halo CTA-pair kernel instance for NT=%d", NT);
}

}  // namespace dbv
