// Resident-halo tcgen05 convolution on CTA PAIRS (cta_group::2, sm_100a) for the stride-1 layers of the fp16 tail of
// DBV_PREC_MIXED (convT6, convT8, head): single-plane fp16 activations x fp16 hi/lo weights.
//
// tc_halo.cu issues, per tap and k-step, ONE MMA  A[128 positions x 16] x [B_hi | B_lo][2*NT x 16]: with M = 128 the fetch of
// the activation operand from shared memory (128 rows x 32 bytes = 32 cycles) is what an MMA costs at small N
// (max(N/2, 32 + N/4) cycles: tools/mma_rate.cu), and these three layers are bound by exactly that (clock64 breakdown:
// the issuing thread is busy 93-97 % of the kernel at 49.6 / 57 / 73 cycles per MMA).  A pair of CTAs (the two SMs of a
// TPC) computes the same band of TWO stamps at once instead: each CTA holds its own halo band (its 128 A rows of the M = 256
// operand) and only HALF of the B operand — the leader the hi weight blocks, its peer the lo blocks, which is precisely how
// cta_group::2 splits the N = 2*NT rows of B between the two shared memories.  Per SM an MMA then fetches 128 A rows + NT
// B rows instead of 128 + 2*NT: by the single-CTA cost model 40 instead of 48 cycles at N = 64, 48 instead of 64 at N = 128.
// Measured per layer: see halo_pair_supported below.
//
//   both CTAs  warp 0     TMA producer: own weight half once, own halo band per item; all bytes of the pair complete on the
//                         LEADER's barriers (cp.async.bulk.tensor ... .cta_group::2)
//   leader     warp 1     MMA issuer: the same flat host-made op table as tc_halo.cu (M = 256 descriptors); tcgen05.commit
//                         multicasts the band release (aempty) and the accumulator hand-over (tfull) to both CTAs
//   both CTAs  warps 2-9  epilogue of the CTA's own stamp (same item table); all 16 warps of the pair release the slot on the
//                         leader's tempty
// Work item g of a launch = band (g / PB) of the stamp pair (g % PB), PB = ceil(B / 2); CTA `rank` takes stamp 2 * pair + rank
// (one past the end for an odd B: its loads are zero-filled or stale, its stores masked).
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"
#include <mutex>
#include <type_traits>

namespace dbv {

constexpr int HALO2_THREADS = 64 + 2 * 128;
constexpr int HALO2_NBUF_MAX = 4;
constexpr int HALO2_NSLOT_MAX = 8;

template <int CBK, int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO2_THREADS, 1) tc_halo2_kernel(const __grid_constant__ HaloLayer L) {
  pdl_trigger();
  constexpr int ROWB = CBK * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sW = sA + L.nbuf * L.buf_bytes;  // this CTA's half of the resident weights: n_wblk blocks of NT x ROWB
  const uint32_t sBar = sW + L.w_bytes + L.tail_pad;
  const uint32_t bar_w = sBar, bar_afull = sBar + 8, bar_aempty = sBar + 40, bar_tfull = sBar + 72, bar_tempty = sBar + 136;
  const uint32_t s_tmem = sBar + 200;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t SW = (uint32_t)L.slot_cols;
  const uint32_t nslot = (512u / SW) < (uint32_t)HALO2_NSLOT_MAX ? (512u / SW) : (uint32_t)HALO2_NSLOT_MAX;
  const uint32_t slot_shift = 31u - (uint32_t)__clz((int)nslot);
  const int units_per_band = L.n_units;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    tma_prefetch_desc(&L.tmB);
    mbar_init(bar_w, 1);
    for (int s = 0; s < HALO2_NBUF_MAX; ++s) {
      mbar_init(bar_afull + 8 * s, 1);   // the leader's copy is the one in use
      mbar_init(bar_aempty + 8 * s, 1);  // multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < HALO2_NSLOT_MAX; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 16);  // 8 epilogue warps of each CTA (leader's copy)
    }
    fence_barrier_init();
  }
  // zeroed slack after every region's TMA box (the slot after the last halo row is the right neighbour of its last pixel)
  for (int r = 0; r < L.nbuf * L.n_regions; ++r)
    for (int i = L.a_box_bytes + 16 * (int)threadIdx.x; i < L.region_bytes; i += 16 * HALO2_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + (r / L.n_regions) * L.buf_bytes + (r % L.n_regions) * L.region_bytes + i), "r"(0) : "memory");
  for (int r = 0; r < L.nbuf; ++r)
    for (int i = L.n_regions * L.region_bytes + 16 * (int)threadIdx.x; i < L.buf_bytes; i += 16 * HALO2_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + r * L.buf_bytes + i), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc2(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const long long PB = (L.B + 1) / 2;                 // stamp pairs
  const long long total = L.total_bands;              // = PB * bands_per_img pair items
  const uint32_t cid = cluster_id_x(), ncl = nclusters_x();
  const long long g0 = total * cid / ncl, g1 = total * (cid + 1) / ncl;
  const long long yb0 = g0 / PB;
  const long long j_first = g0 - yb0 * PB;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t w0 = map_to_rank(bar_w, 0), afull0 = map_to_rank(bar_afull, 0);
      if (rank == 0) mbar_expect_tx(bar_w, 2u * (uint32_t)(L.n_wblk * NT * ROWB));
      for (int blk = 0; blk < L.n_wblk; ++blk)  // the leader keeps the hi block of every (tap, chunk), its peer the lo block
        tma2_load_2d(sW + blk * (NT * ROWB), &L.tmB, w0, 0, ((int)L.w_src[blk] + (int)rank) * L.w_rows_per_blk);
      int stage = 0;
      uint32_t phase = 0;
      long long j = j_first;
      int y0 = (int)yb0 * L.R;
      for (long long g = g0; g < g1; ++g) {
        mbar_wait_cluster(bar_aempty + 8 * stage, phase ^ 1u);
        if (rank == 0) mbar_expect_tx(bar_afull + 8 * stage, 2u * (uint32_t)(L.n_regions * L.a_box_bytes));
        const int b = (int)(2 * j + rank);
        for (int r = 0; r < L.n_regions; ++r)
          tma2_load_5d(sA + stage * L.buf_bytes + r * L.region_bytes, &L.tmA, afull0 + 8 * stage, L.region_coff[r], -L.pad, y0 - L.pad_top, L.region_c3[r], b);
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
        if (++j == PB) { j = 0; y0 += L.R; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      mbar_wait_cluster(bar_w, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t u = 0;
      constexpr uint32_t HI = smem_desc_hi<ROWB>();
      const uint32_t w16 = kSmemDescLoConst | (sW >> 4);
      for (long long g = g0; g < g1; ++g) {
        mbar_wait_cluster(bar_afull + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a16 = kSmemDescLoConst | ((sA + stage * L.buf_bytes) >> 4);
        int i = 0;
        for (int k = 0; k < units_per_band; ++k, ++u) {
          const uint32_t slot = u & (nslot - 1);
          mbar_wait_cluster(bar_tempty + 8 * slot, ((u >> slot_shift) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d0 = tmem_base + slot * SW;
          const int iend = L.unit_op_end[k];
#pragma unroll 4
          for (; i < iend; ++i) {
            const HaloOp e = L.ops[i];
            umma2_f16(d0 + (e.d & 0xffffu), desc64(HI, a16 + e.a), desc64(HI, w16 + e.b), e.idesc, e.d >> 16);
          }
          umma2_commit_mc(bar_tfull + 8 * slot);
        }
        umma2_commit_mc(bar_aempty + 8 * stage);
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    constexpr int NV = (NT % 32 != 0) ? 16 : 32;
    static_assert(NT % NV == 0, "items cover whole channel chunks");
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tempty0 = map_to_rank(bar_tempty, 0);
    const bool has_alpha = L.o.alpha != nullptr;
    const float4* alpha4 = reinterpret_cast<const float4*>(L.o.alpha);
    const bool le1 = L.o.alpha_le1 != 0;
    const uint32_t npix = (uint32_t)(L.o.OH * L.o.OW);
    auto run = [&](auto MODE, auto PLANES, auto F16) {
      OutSpec o = L.o;
      if constexpr (decltype(MODE)::value >= 0) {
        o.mode = decltype(MODE)::value;
        o.planes = decltype(PLANES)::value;
        o.f16 = decltype(F16)::value;
      }
      uint32_t u = 0;
      long long j = j_first;
      int y0 = (int)yb0 * L.R;
      for (long long g = g0; g < g1; ++g) {
        const long long b = 2 * j + rank;
        const bool b_ok = b < L.B;
        for (int k = 0; k < units_per_band; ++k, ++u) {
          const uint32_t slot = u & (nslot - 1);
          bool waited = false;
          const int it0 = k ? (int)L.unit_item_end[k - 1] : 0, it1 = (int)L.unit_item_end[k];
#pragma unroll 1
          for (int item = it0 + grp; item < it1; item += 2) {
            const uint32_t it = L.items[item];
            const int m = (int)((it >> 11) & 31u), q = (int)((it >> 16) & 7u);
            const int p = 128 * m + row;
            const int ly = (int)__umulhi((uint32_t)p, L.magic_wp), sx = p - ly * L.WP, sy = y0 + ly;
            const bool ok = b_ok && ly < L.R && sx < L.W && sy < L.H;
            const int oy = sy, ox = sx;  // stride-1 layers only: one output class
            const uint32_t tcol = lane_base + slot * SW + (it & 511u);
            const int c0 = q * NV;
            float4 al[NV / 4];
            if (has_alpha && ok) {
              const uint32_t off = (uint32_t)(c0 >> 2) * npix + (uint32_t)(oy * L.o.OW + ox);
#pragma unroll
              for (int jj = 0; jj < NV / 4; ++jj) al[jj] = __ldg(alpha4 + off + (uint32_t)jj * npix);
            }
            if (!waited) {
              mbar_wait_cluster_relaxed(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
              tc_fence_after();
              waited = true;
            }
            float v[NV], w[NV];
            tmem_ld_issue<NV>(tcol, v);
            tmem_ld_issue<NV>(tcol + (uint32_t)NT, w);  // + the A x B_lo partial product (second half of the tile's columns)
            tmem_ld_wait<NV>(v);
            tmem_ld_wait<NV>(w);
#pragma unroll
            for (int jj = 0; jj < NV; jj += 2) add2(v[jj], v[jj + 1], w[jj], w[jj + 1]);
            if (ok) {
#pragma unroll
              for (int jj = 0; jj < NV; ++jj) v[jj] += L.bias_c[c0 + jj];
              if (has_alpha) {
#pragma unroll
                for (int jj = 0; jj < NV / 4; ++jj) prelu4(v[4 * jj + 0], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3], al[jj], le1);
              } else if (L.o.relu) {
#pragma unroll
                for (int jj = 0; jj < NV; ++jj) v[jj] = fmaxf(v[jj], 0.f);
              }
              store_act<NV>(o, b, oy, ox, c0, v);
            }
          }
          if (!waited) {
            mbar_wait_cluster_relaxed(bar_tfull + 8 * slot, (u >> slot_shift) & 1u);
            tc_fence_after();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + 8 * slot);
        }
        if (++j == PB) { j = 0; y0 += L.R; }
      }
    };
    using std::integral_constant;
    const int om = L.o.mode, op = L.o.planes, of = L.o.f16;
    if (om == OUT_BF16_NHWC && op == 1 && of == 1) run(integral_constant<int, OUT_BF16_NHWC>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{});
    else if (om == OUT_HEAD) run(integral_constant<int, OUT_HEAD>{}, integral_constant<int, 1>{}, integral_constant<int, 0>{});
    else run(integral_constant<int, -1>{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int CBK, int NT>
static int launch_halo2_one(const HaloLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_halo2_kernel<CBK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_MAX_SMEM); });
  if (attr_err != cudaSuccess) return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_halo2_kernel<%d,%d>): %s", CBK, NT, cudaGetErrorString(attr_err));
  if (L.total_bands <= 0) return DBV_OK;
  const long long clusters = L.total_bands < max_ctas / 2 ? L.total_bands : max_ctas / 2;
  launch_pdl(tc_halo2_kernel<CBK, NT>, (unsigned)(2 * clusters), HALO2_THREADS, L.smem_bytes, st, L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

// Measured on B200 (plan tuner, 2368 stamps; single CTA -> pair): convT6 (N = 2*NT = 128) 0.250 -> 0.207 ms, convT8 (N = 64)
// 0.393 -> 0.376 ms, head (N = 32) 0.288 -> 0.262 ms.  The first version handed the accumulator slots back with
// mbarrier.arrive.release.cluster and was SLOWER than the single-CTA kernel for convT8 and the head (0.457 / 0.317 ms): a
// release at cluster scope makes every epilogue warp wait until the activations it has just stored are visible to the other
// SM, once per unit.  The hand-over publishes nothing through memory, so it is now a relaxed arrive (tc_pair_ptx.cuh); the
// same change in tc_pair.cu / tc_pairh.cu bought 3-7 % on conv5 / convT5.  tools/mma2_rate.cu: a pair MMA costs 39 / 43 / 64 /
// 128 cycles at N = 32 / 64 / 128 / 256 for twice the rows of a single-CTA MMA (40 / 48 / 64 / 128).
bool halo_pair_supported(int CBK, int NT) { return (CBK == 64 && NT == 64) || (CBK == 32 && (NT == 32 || NT == 16)); }

int launch_halo_pair_layer(const HaloLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
  if (CBK == 64 && NT == 64) return launch_halo2_one<64, 64>(L, max_ctas, st);
  if (CBK == 32 && NT == 32) return launch_halo2_one<32, 32>(L, max_ctas, st);
  if (CBK == 32 && NT == 16) return launch_halo2_one<32, 16>(L, max_ctas, st);
  return fail(DBV_ERR_UNSUPPORTED, "no CTA-pair halo kernel instance for CBK=%d NT=%d", CBK, NT);
}

}  // namespace dbv
