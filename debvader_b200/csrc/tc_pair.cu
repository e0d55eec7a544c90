// tcgen05 implicit-GEMM convolution on CTA PAIRS (cta_group::2, sm_100a) for the many-channel layers.
//
// The N >= 128 layers are bound by moving the WEIGHT operand: with one CTA per 128-row tile every SM streams
// the whole packed weight tensor of the layer through L2 -> shared memory once per tile (measured: conv 256->256
// at 8x8 moved 7.2 GB per 4096 stamps, 0.75 ms against an MMA floor of 0.41 ms).  A cluster of two CTAs
// (the two SMs of a TPC) computes a 256-row tile pair instead: each CTA loads its own 128 activation rows and
// only HALF of every weight box (N/2 rows); one thread of the leader CTA issues tcgen05.mma.cta_group::2
// (M = 256), and the hardware reads the two halves of B from the two CTAs' shared memories.  Weight traffic
// per SM (L2 -> smem fill and smem -> tensor-core fetch) halves.
//
//   both CTAs   warp 0   TMA producer: A_hi, A_lo boxes of the CTA's own tile, B_hi/2, B_lo/2 boxes; the bytes of BOTH
//                        CTAs complete on the LEADER's `full` barrier (cp.async.bulk.tensor ... .cta_group::2)
//   leader      warp 1   MMA issuer: D += A_hi B_hi + A_hi B_lo + A_lo B_hi, M=256; tcgen05.commit multicasts the
//                        stage release (`empty`) and the accumulator hand-over (`tfull`) to both CTAs
//   both CTAs   warps 2-9 epilogue of the CTA's own 128 rows (same code as tc_conv.cu); the 8 warps of the pair arrive
//                        on the leader's `tempty`
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"
#include <mutex>

namespace dbv {

constexpr int TP_THREADS = 64 + 2 * 128;
constexpr int TP_MAX_SMEM = 232448;

template <int CBK, int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1) tc_pair_kernel(const __grid_constant__ TcLayer L) {
  pdl_trigger();
  constexpr int ROWB = CBK * 2;
  constexpr int A_BYTES = 128 * ROWB;
  constexpr int BH_BYTES = (NT / 2) * ROWB;  // this CTA's half of one weight box
  // M = 256 (both CTAs), N = NT
  const uint32_t IDESC = (1u << 4) | idesc_ab_fmt(L.ab_f16) | ((uint32_t)(NT >> 3) << 17) | ((256u >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int STAGES = L.stages;
  const uint32_t stage_bytes = (uint32_t)L.stage_bytes;
  const int parts = L.x3 ? 2 : 1;
  const uint32_t offB = (uint32_t)parts * A_BYTES;
  const uint32_t sBar = base + (uint32_t)STAGES * stage_bytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 64, bar_tfull = sBar + 128, bar_tempty = sBar + 144, s_tmem = sBar + 160;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr uint32_t SLOTW = (uint32_t)tmem_cols_for(NT);  // two accumulator slots

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    tma_prefetch_desc(&L.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);   // leader's copy is the one in use
      mbar_init(bar_empty + 8 * s, 1);  // multicast commit from the leader's MMA warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 16);  // 8 epilogue warps of each CTA (leader's copy is the one in use)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // nothing produced or consumed by the previous kernel is touched above this line

  const int tiles_img = L.tiles_x * L.tiles_y;
  const long long items = L.pair_items;  // clusters' work items: n_cls x ceil(m tiles / 2) x n_tiles_n
  const long long items_per_cls = items / L.n_cls;
  const uint32_t cid = cluster_id_x(), ncl = nclusters_x();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = 2u * (uint32_t)(parts * (L.a_bytes + BH_BYTES));  // both CTAs' bytes land on the leader's barrier
      const uint32_t full0 = map_to_rank(bar_full, 0);
      for (long long w = cid; w < items; w += ncl) {
        const int c = (int)(w / items_per_cls);
        long long r = w - (long long)c * items_per_cls;
        const int nt = (int)(r % L.n_tiles_n);
        const long long m = 2 * (r / L.n_tiles_n) + rank;  // this CTA's m tile (may be one past the end: all-zero loads, masked stores)
        const int ti = (int)(m % tiles_img);
        const int bt = (int)(m / tiles_img);
        const int x0 = (ti % L.tiles_x) * L.TW, y0 = (ti / L.tiles_x) * L.TH, b0 = bt * L.TB;
        const TcClass cl = L.cls[c];
        for (int kb = 0; kb < cl.nkb; ++kb) {
          const TcKBlock K = L.kb[cl.kb_begin + kb];
          const uint32_t sS = base + (uint32_t)stage * stage_bytes, bar = full0 + 8 * stage;
          mbar_wait_cluster(bar_empty + 8 * stage, phase ^ 1u);
          if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, tx);
          const int brow = K.b_row + nt * NT + (int)rank * (NT / 2);
          tma2_load_5d(sS, &L.tmA, bar, K.c_off, x0 + K.dx, y0 + K.dy, K.plane, b0);
          tma2_load_2d(sS + offB, &L.tmB, bar, 0, brow);
          if (L.x3) {
            tma2_load_5d(sS + A_BYTES, &L.tmA, bar, K.c_off + L.lo_coff, x0 + K.dx, y0 + K.dy, K.plane, b0);
            tma2_load_2d(sS + offB + BH_BYTES, &L.tmB, bar, 0, brow + L.lo_brow);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;
      constexpr uint32_t HI = smem_desc_hi<ROWB>();
      for (long long w = cid; w < items; w += ncl, ++u) {
        const int c = (int)(w / items_per_cls);
        const int nkb = L.cls[c].nkb;
        const uint32_t slot = u & 1u;
        mbar_wait_cluster(bar_tempty + 8 * slot, ((u >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * SLOTW;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_cluster(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sS = base + (uint32_t)stage * stage_bytes;
          const uint32_t ahi = kSmemDescLoConst | ((sS & 0x3FFFFu) >> 4);
          const uint32_t alo = ahi + (A_BYTES >> 4);
          const uint32_t bhi = kSmemDescLoConst | (((sS + offB) & 0x3FFFFu) >> 4);
          const uint32_t blo = bhi + (BH_BYTES >> 4);
          if (L.x3) {
#pragma unroll
            for (int k = 0; k < CBK / 16; ++k) {
              umma2_f16(d_tmem, desc64(HI, ahi + 2 * k), desc64(HI, bhi + 2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
              umma2_f16(d_tmem, desc64(HI, ahi + 2 * k), desc64(HI, blo + 2 * k), IDESC, 1u);
              umma2_f16(d_tmem, desc64(HI, alo + 2 * k), desc64(HI, bhi + 2 * k), IDESC, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < CBK / 16; ++k)
              umma2_f16(d_tmem, desc64(HI, ahi + 2 * k), desc64(HI, bhi + 2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_mc(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma2_commit_mc(bar_tfull + 8 * slot);
      }
    }
  } else {
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int rows_img = L.TW * L.TH;
    const int tb = row / rows_img;
    const int rr = row - tb * rows_img;
    const int ty = rr / L.TW, tx = rr - ty * L.TW;
    const bool row_ok = tb < L.TB;
    constexpr int NV = 32;
    constexpr int NCHK = NT / NV;
    static_assert(NT % 32 == 0, "pair kernel: NT must be a multiple of 32");
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tempty0 = map_to_rank(bar_tempty, 0);
    uint32_t u = 0;
    for (long long w = cid; w < items; w += ncl, ++u) {
      const int c = (int)(w / items_per_cls);
      long long r = w - (long long)c * items_per_cls;
      const int nt = (int)(r % L.n_tiles_n);
      const long long m = 2 * (r / L.n_tiles_n) + rank;
      const int ti = (int)(m % tiles_img);
      const int bt = (int)(m / tiles_img);
      const int sx = (ti % L.tiles_x) * L.TW + tx, sy = (ti / L.tiles_x) * L.TH + ty;
      const long long b = (long long)bt * L.TB + tb;
      const TcClass cl = L.cls[c];
      const bool ok = row_ok && b < L.B && sx < L.SW && sy < L.SH;
      int oy = cl.oy0 + cl.osy * sy, ox = cl.ox0 + cl.osx * sx;
      int cbase = nt * NT, boff = 0;
      if (L.nt_pixel_mode) {
        oy = nt / L.o.OW;
        ox = nt - oy * L.o.OW;
        cbase = 0;
        boff = nt * NT;
      }
      const uint32_t slot = u & 1u;
      const uint32_t tcol = lane_base + slot * SLOTW;
      // the two groups split the tile's 32-channel chunks (group g takes chunks q = g, g+2, ...)
      ActRegs<NV> ra;
      if (grp < NCHK) act_prefetch<NV>(L.o, ok, oy, ox, cbase + grp * NV, boff, ra);
      mbar_wait_cluster_relaxed(bar_tfull + 8 * slot, (u >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int q = grp; q < NCHK; q += 2) {
        if (q != grp) act_prefetch<NV>(L.o, ok, oy, ox, cbase + q * NV, boff, ra);
        float v[NV];
        tmem_ld_issue<NV>(tcol + (uint32_t)(q * NV), v);
        tmem_ld_wait<NV>(v);
        if (ok) {
          act_apply<NV>(L.o, oy, ox, cbase + q * NV, boff, ra, v);
          store_act<NV>(L.o, b, oy, ox, cbase + q * NV, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + 8 * slot);  // nothing to publish through memory: see tc_pair_ptx.cuh
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its shared memory until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int CBK, int NT>
static int launch_pair_one(const TcLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_pair_kernel<CBK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_MAX_SMEM);
  });
  if (attr_err != cudaSuccess)
    return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_pair_kernel<%d,%d>): %s", CBK, NT, cudaGetErrorString(attr_err));
  if (L.pair_items <= 0) return DBV_OK;
  long long clusters = L.pair_items < max_ctas / 2 ? L.pair_items : max_ctas / 2;
  const int smem = L.stages * L.stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/;
  if (L.stages < 2 || L.stages > 8 || smem > TP_MAX_SMEM)
    return fail(DBV_ERR_STATE, "tc_pair_kernel<%d,%d>: bad stage plan (%d x %d B)", CBK, NT, L.stages, L.stage_bytes);
  launch_pdl(tc_pair_kernel<CBK, NT>, (unsigned)(2 * clusters), TP_THREADS, smem, st, L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

// stage plan of the pair kernel: A_hi[, A_lo], B_hi/2[, B_lo/2]
void tc_pair_stage_plan(TcLayer& L, int CBK, int NT) {
  const int parts = L.x3 ? 2 : 1;
  L.stage_bytes = parts * (128 * CBK * 2 + (NT / 2) * CBK * 2);
  int st = (TP_MAX_SMEM - 1536) / L.stage_bytes;
  L.stages = st > 8 ? 8 : st;
  L.wide = 0;
}

bool tc_pair_supported(int CBK, int NT) { return CBK == 64 && (NT == 64 || NT == 128 || NT == 256); }

int launch_tc_pair(const TcLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
  if (CBK == 64 && NT == 64) return launch_pair_one<64, 64>(L, max_ctas, st);
  if (CBK == 64 && NT == 128) return launch_pair_one<64, 128>(L, max_ctas, st);
  if (CBK == 64 && NT == 256) return launch_pair_one<64, 256>(L, max_ctas, st);
  return fail(DBV_ERR_UNSUPPORTED, "no CTA-pair kernel instance for CBK=%d NT=%d", CBK, NT);
}

}  // namespace dbv
