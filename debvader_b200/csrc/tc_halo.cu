// tcgen05 implicit-GEMM convolution with a RESIDENT HALO TILE (sm_100a).
//
// tc_conv.cu re-reads the activation tile from L2 once per 3x3 tap (and, in the bf16x3 split, once
// per hi/lo pairing): 27 x 10 KB per 128 output pixels for a 32-channel layer, which made the
// large-image / few-channel layers L2-bandwidth bound (profiles/r01_v1_summary.md).  Here a CTA loads
// a band of R+2 input rows x (W+2) columns ONCE per plane (TMA, out-of-bounds zero fill = TF "SAME"
// padding) and all taps, all output-parity classes of a stride-2 transposed conv and all hi/lo
// pairings read it as *row-shifted windows of the same shared-memory tile*: output position
// p = y*(W+2)+x of the band, tap (dy,dx) -> smem row p + (dy+1)*(W+2) + (dx+1).  That works because
// tcgen05 applies the 64B/128B swizzle to absolute shared-memory address bits, so a descriptor
// whose start is shifted by whole rows stays consistent with what TMA wrote (measured:
// tests/test_gpu_network.py::test_probe_descriptor_row_shift).  Weights are loaded once per CTA and
// stay resident.  Two columns per row (x = W, W+1) compute garbage that the epilogue drops.
//
//   warp 0    TMA producer (weights once; one halo band per work item, ring of nbuf buffers)
//   warp 1    MMA issuer: for class, for 128-position tile, for (tap, chunk, pairing), CBK/16 MMAs
//   warps 2-9 epilogue, two groups of 4 (shared with tc_conv.cu: bias, PReLU(h,w,c), ReLU/crop/split, bf16 hi/lo)
#include "tc_ptx.cuh"
#include <mutex>

namespace dbv {

constexpr int EPI_SUBGROUPS = 1;  // epilogue groups (of 4 warps) per accumulator buffer; 2 was measured slower (L1-bound, spills)
constexpr int HALO_THREADS = 64 + 2 * EPI_SUBGROUPS * 128;  // TMA warp, MMA warp, epilogue warps
constexpr int HALO_TBUF_COLS = 256;  // TMEM columns per accumulator buffer (2 buffers)

// NOSWZ (encoder conv1, Cin = 6 padded to 8): one 16-byte row per pixel, no swizzle.  A K=16 MMA operand is
// then TWO ADJACENT PIXELS: core-matrix stride along K (LBO) = 16 bytes = the pixel pitch, so the 3x3x8
// im2col never exists anywhere — the (kx, channel) axis of each kernel row is read as overlapping
// windows of the halo tile.  Per kernel row ky: pixels (x-1, x) and (x+1, x+2[zero weights]).
template <int CBK, int NT, bool NOSWZ>
__global__ void __launch_bounds__(HALO_THREADS, 1) tc_halo_kernel(const __grid_constant__ HaloLayer L) {
  constexpr int ROWB = NOSWZ ? 16 : CBK * 2;
  constexpr int KSTEPS = NOSWZ ? 1 : CBK / 16;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * NT) >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base;                              // resident weights: n_wblk blocks of NT x ROWB
  const uint32_t sA = base + L.w_bytes;                  // nbuf x n_regions x region_bytes
  const uint32_t sBar = sA + L.nbuf * L.n_regions * L.region_bytes + L.tail_pad;
  const uint32_t bar_w = sBar, bar_afull = sBar + 8, bar_aempty = sBar + 24, bar_tfull = sBar + 40, bar_tempty = sBar + 56;
  const uint32_t s_tmem = sBar + 72;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    if (!NOSWZ) tma_prefetch_desc(&L.tmB);
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4 * EPI_SUBGROUPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long total = L.total_bands;

  if (warp == 0) {
    if (elect_one()) {
      if constexpr (NOSWZ) {
        mbar_expect_tx(bar_w, (uint32_t)L.w_bytes);
        bulk_load(sW, L.w_img, (uint32_t)L.w_bytes, bar_w);
      } else {
        mbar_expect_tx(bar_w, (uint32_t)(L.n_wblk * NT * ROWB));
        for (int blk = 0; blk < L.n_wblk; ++blk) tma_load_2d(sW + blk * (NT * ROWB), &L.tmB, bar_w, 0, blk * L.w_rows_per_blk);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const long long b = t / L.bands_per_img;
        const int y0 = (int)(t - b * L.bands_per_img) * L.R;
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        mbar_expect_tx(bar_afull + 8 * stage, (uint32_t)(L.n_regions * L.a_box_bytes));
        for (int r = 0; r < L.n_regions; ++r)
          tma_load_5d(sA + (stage * L.n_regions + r) * L.region_bytes, &L.tmA, bar_afull + 8 * stage, L.region_coff[r], -L.pad, y0 - L.pad, 0, (int)b);
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(bar_w, 0);
      int stage = 0;
      uint32_t phase = 0;
      int tb = 0;
      uint32_t tphase = 0;
      // descriptor words: swizzled K-major rows, or (NOSWZ) interleaved 8x16-byte core matrices with
      // A: SBO 128 B (8 pixels), LBO 16 B (next pixel);  B: SBO 256 B, LBO 128 B (host-packed image)
      constexpr uint32_t HI = NOSWZ ? ((128u >> 4) | (1u << 14)) : smem_desc_hi<ROWB>();
      constexpr uint32_t HIB = NOSWZ ? ((256u >> 4) | (1u << 14)) : smem_desc_hi<ROWB>();
      constexpr uint32_t LOB = NOSWZ ? ((128u >> 4) << 16) : kSmemDescLoConst;
      constexpr uint32_t MSTEP = (128 * ROWB) >> 4;
      const uint32_t w16 = LOB | (sW >> 4);
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        mbar_wait(bar_tempty + 8 * tb, tphase ^ 1u);
        mbar_wait(bar_afull + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a16 = kSmemDescLoConst | ((sA + stage * L.n_regions * L.region_bytes) >> 4);
        for (int c = 0; c < ((L.dbg_skip & 1) ? 0 : L.n_cls); ++c) {
          const int kb0 = L.cls[c].kb_begin, nkb = L.cls[c].nkb;
          const uint32_t DW = (uint32_t)(L.wide ? 2 * NT : NT);  // accumulator columns per tile
          const uint32_t d0 = tmem_base + (uint32_t)(tb * HALO_TBUF_COLS) + (uint32_t)(c * L.ntiles) * DW;
          // k-block outer, tile inner: the per-k-block table lookup is amortised over ntiles * CBK/16 MMAs
          for (int kb = 0; kb < nkb; ++kb) {
            const TcKBlock K = L.kb[kb0 + kb];
            uint32_t alo = a16 + (uint32_t)(uint16_t)K.c_off;
            const uint32_t blo = w16 + (uint32_t)K.b_row;
            const uint32_t idesc = K.dy ? IDESC2 : IDESC;
            uint32_t d = d0;
            for (int m = 0; m < L.ntiles; ++m, alo += MSTEP, d += DW) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) umma_f16(d, desc64(HI, alo + 2 * k), desc64(HIB, blo + 2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(bar_aempty + 8 * stage);
        umma_commit(bar_tfull + 8 * tb);
        if (++stage == L.nbuf) { stage = 0; phase ^= 1u; }
        tb ^= 1;
        if (tb == 0) tphase ^= 1u;
      }
    }
  } else {
    // 16 epilogue warps = 4 groups of 4 (one warp per TMEM lane quadrant in each group).  Groups 0,1 drain
    // accumulator buffer 0, groups 2,3 buffer 1; within a band the two groups take alternate items
    // (tile, 32-channel chunk).  The epilogue is dependent-issue bound (ncu: ~0.2 IPC per warp), so it is
    // the number of resident warps per scheduler — 4 — that hides its latency.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int half = grp / EPI_SUBGROUPS, sub = grp % EPI_SUBGROUPS;
    const int row = quad * 32 + lane;
    constexpr int NV = (NT % 32 == 0) ? 32 : 16;
    constexpr int NCHK = NT / NV;
    const int n_items = L.n_cls * L.ntiles * NCHK;
    const uint32_t DW = (uint32_t)(L.wide ? 2 * NT : NT);
    int tb = 0;
    uint32_t tphase = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if (tb == half) {
        const long long b = t / L.bands_per_img;
        const int y0 = (int)(t - b * L.bands_per_img) * L.R;
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tb * HALO_TBUF_COLS);
        mbar_wait(bar_tfull + 8 * tb, tphase);
        tc_fence_after();
#pragma unroll 1
        for (int q = sub; q < ((L.dbg_skip & 2) ? 0 : n_items); q += EPI_SUBGROUPS) {
          const int tm = q / NCHK, c0 = (q - tm * NCHK) * NV;
          const int c = tm / L.ntiles, m = tm - c * L.ntiles;
          const int p = 128 * m + row;
          const int ly = p / L.WP, sx = p - ly * L.WP, sy = y0 + ly;
          const bool ok = ly < L.R && sx < L.W && sy < L.H;
          const int oy = L.cls[c].oy0 + L.cls[c].osy * sy, ox = L.cls[c].ox0 + L.cls[c].osx * sx;
          ActRegs<NV> ra;
          act_prefetch<NV>(L.o, ok, oy, ox, c0, 0, ra);
          float v[NV];
          tmem_ld<NV>(tbase + (uint32_t)tm * DW + (uint32_t)c0, v);
          if (L.wide) {  // + the A_hi x B_lo partial product held in the second half of the tile's columns
            float w[NV];
            tmem_ld<NV>(tbase + (uint32_t)tm * DW + (uint32_t)(NT + c0), w);
#pragma unroll
            for (int j = 0; j < NV; ++j) v[j] += w[j];
          }
          if (ok) {
            act_apply<NV>(L.o, oy, ox, c0, 0, ra, v);
            store_act<NV>(L.o, b, oy, ox, c0, v);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * tb);
      }
      tb ^= 1;
      if (tb == 0) tphase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int CBK, int NT, bool NOSWZ = false>
static int launch_halo_one(const HaloLayer& L, int max_ctas, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_halo_kernel<CBK, NT, NOSWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_MAX_SMEM);
  });
  if (attr_err != cudaSuccess)
    return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_halo_kernel<%d,%d>): %s", CBK, NT, cudaGetErrorString(attr_err));
  const long long grid = L.total_bands < max_ctas ? L.total_bands : max_ctas;
  if (grid <= 0) return DBV_OK;
  tc_halo_kernel<CBK, NT, NOSWZ><<<(unsigned)grid, HALO_THREADS, L.smem_bytes, st>>>(L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

bool halo_layer_supported(int CBK, int NT) {
  if (CBK == 16) return NT == 32;  // encoder conv1, no-swizzle mode
  if (CBK == 32) return NT == 16 || NT == 32 || NT == 64;
  if (CBK == 64) return NT == 32 || NT == 64 || NT == 128;
  return false;
}

int launch_halo_layer(const HaloLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
#define DBV_HALO_CASE(cb, nt) \
  if (CBK == cb && NT == nt) return launch_halo_one<cb, nt>(L, max_ctas, st);
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
