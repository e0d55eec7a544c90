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

constexpr int TC_EPI_SUBGROUPS = 1;  // epilogue groups (of 4 warps) per accumulator buffer; 2 was measured slower (L1-bound, spills)
constexpr int TC_THREADS = 64 + 2 * TC_EPI_SUBGROUPS * 128;  // TMA warp, MMA warp, epilogue warps

template <int CBK, int NT>
struct TcCfg {
  static constexpr int ROWB = CBK * 2;
  static constexpr int A_STAGE = 128 * ROWB;
  static constexpr int B_STAGE = NT * ROWB;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int STAGES = (196608 / STAGE) > 8 ? 8 : (196608 / STAGE);
  static constexpr int TMEM_COLS = tmem_cols_for(2 * NT);
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=BF16 [7,10)/[10,13),
  // K-major both, N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
};

template <int CBK, int NT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ TcLayer L) {
  using Cfg = TcCfg<CBK, NT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * Cfg::A_STAGE;
  const uint32_t sBar = sB + STAGES * Cfg::B_STAGE;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_tfull = sBar + 16 * STAGES, bar_tempty = bar_tfull + 16;
  const uint32_t s_tmem = bar_tempty + 16;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (s_tmem - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.tmA);
    tma_prefetch_desc(&L.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4 * TC_EPI_SUBGROUPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long total = L.total_tiles;
  const int tiles_img = L.tiles_x * L.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
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
          mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          mbar_expect_tx(bar_full + 8 * stage, (uint32_t)(L.a_bytes + L.b_bytes));
          tma_load_5d(sA + stage * Cfg::A_STAGE, &L.tmA, bar_full + 8 * stage, K.c_off, x0 + K.dx, y0 + K.dy, K.plane, b0 - L.dbg_shift_rows);
          tma_load_2d(sB + stage * Cfg::B_STAGE, &L.tmB, bar_full + 8 * stage, 0, K.b_row + nt * NT);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int c = (int)(t / L.tiles_per_cls);
        const int nkb = L.cls[c].nkb;
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * NT);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          uint64_t adesc = make_smem_desc<Cfg::ROWB>(sA + stage * Cfg::A_STAGE + L.dbg_shift_rows * Cfg::ROWB);
          if (L.dbg_base_mode == 1)
            adesc |= (uint64_t)(((sA + stage * Cfg::A_STAGE + L.dbg_shift_rows * Cfg::ROWB) >> 7) & 7u) << 49;
          const uint64_t bdesc = make_smem_desc<Cfg::ROWB>(sB + stage * Cfg::B_STAGE);
#pragma unroll
          for (int k = 0; k < CBK / 16; ++k)
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), Cfg::IDESC, (kb | k) != 0 ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * as);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // 16 epilogue warps = 4 groups of 4 (one warp per TMEM lane quadrant in each group).  Groups 0,1 drain
    // accumulator buffer 0 (even tiles), groups 2,3 buffer 1; the two groups of a buffer take alternate
    // 32-channel chunks.  4 resident epilogue warps per scheduler hide the dependent-issue latency.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int half = grp / TC_EPI_SUBGROUPS, sub = grp % TC_EPI_SUBGROUPS;
    const int row = quad * 32 + lane;
    const int rows_img = L.TW * L.TH;
    const int tb = row / rows_img;
    const int rr = row - tb * rows_img;
    const int ty = rr / L.TW, tx = rr - ty * L.TW;
    const bool row_ok = tb < L.TB;
    constexpr int NV = (NT % 32 == 0) ? 32 : 16;
    constexpr int NCHK = NT / NV;
    int as = 0;
    uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      if (as == half) {
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
        if (L.nt_pixel_mode) {
          oy = nt / L.o.OW;
          ox = nt - oy * L.o.OW;
          cbase = 0;
          boff = nt * NT;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * NT);
        ActRegs<NV> ra;
        if (sub < NCHK) act_prefetch<NV>(L.o, ok, oy, ox, cbase + sub * NV, boff, ra);
        mbar_wait(bar_tfull + 8 * as, aphase);
        tc_fence_after();
#pragma unroll 1
        for (int q = sub; q < NCHK; q += TC_EPI_SUBGROUPS) {
          if (q != sub) act_prefetch<NV>(L.o, ok, oy, ox, cbase + q * NV, boff, ra);
          float v[NV];
          tmem_ld<NV>(taddr + q * NV, v);
          if (ok) {
            act_apply<NV>(L.o, oy, ox, cbase + q * NV, boff, ra, v);
            store_act<NV>(L.o, b, oy, ox, cbase + q * NV, v);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int CBK, int NT>
static int launch_one(const TcLayer& L, int max_ctas, cudaStream_t st) {
  using Cfg = TcCfg<CBK, NT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_conv_kernel<CBK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  if (attr_err != cudaSuccess)
    return fail(DBV_ERR_CUDA, "cudaFuncSetAttribute(tc_conv_kernel<%d,%d>, smem=%d): %s", CBK, NT, Cfg::SMEM,
                cudaGetErrorString(attr_err));
  long long grid = L.total_tiles < max_ctas ? L.total_tiles : max_ctas;
  if (grid <= 0) return DBV_OK;
  tc_conv_kernel<CBK, NT><<<(unsigned)grid, TC_THREADS, Cfg::SMEM, st>>>(L);
  DBV_LAUNCH_CHECK();
  return DBV_OK;
}

bool tc_layer_supported(int CBK, int NT) {
  if (CBK == 32) return NT == 16 || NT == 32 || NT == 64;
  if (CBK == 64) return NT == 32 || NT == 64 || NT == 112 || NT == 128 || NT == 256;
  return false;
}

int launch_tc_layer(const TcLayer& L, int CBK, int NT, int max_ctas, cudaStream_t st) {
#define DBV_TC_CASE(cb, nt) \
  if (CBK == cb && NT == nt) return launch_one<cb, nt>(L, max_ctas, st);
  DBV_TC_CASE(32, 16)
  DBV_TC_CASE(32, 32)
  DBV_TC_CASE(32, 64)
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
                const uint32_t* box, int swizzle_bytes) {
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
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, bx, es,
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
