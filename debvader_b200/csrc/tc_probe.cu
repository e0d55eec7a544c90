// Self-test entry point for the tcgen05/TMA conventions the library relies on: a plain GEMM
//   out[M][N] (fp32) = A[M][K] (bf16, K contiguous) x Bp (bf16, packed [K/CBK][N][CBK])
// run through the SAME kernel instance (tc_conv_kernel<CBK,N>) the network layers use, with a
// dense-layer plan (1x1 tile space, 128 rows of the batch dimension per tile).
//   which = 0 : CBK = 64 (128-byte rows, SWIZZLE_128B)
//   which = 1 : CBK = 32 ( 64-byte rows, SWIZZLE_64B)
#include "kernels.h"
#include "../../include/debvader_b200_debug.h"
#include <vector>

using namespace dbv;

extern "C" int dbv_probe(int which, const void* a_dev, const void* b_dev, float* out_dev, int M, int N, int K, void* stream) {
  DBV_REQUIRE(a_dev && b_dev && out_dev, "dbv_probe: null pointer");
  // which: bits 0-7 = 0 (CBK 64) | 1 (CBK 32); bits 8-15 = A row shift; bits 16-19 = base_offset mode
  const int shift = (which >> 8) & 0xff, base_mode = (which >> 16) & 0xf;
  which &= 0xff;
  const int CBK = which == 0 ? 64 : 32;
  DBV_REQUIRE(which == 0 || which == 1, "dbv_probe: unknown probe %d", which);
  DBV_REQUIRE(M > 0 && K > 0 && K % CBK == 0 && K / CBK <= TC_MAX_KB, "dbv_probe: bad K=%d for CBK=%d", K, CBK);
  DBV_REQUIRE(tc_layer_supported(CBK, N), "dbv_probe: no kernel instance for CBK=%d N=%d", CBK, N);
  static float* zero_bias = nullptr;
  if (!zero_bias) {
    DBV_CUDA(cudaMalloc(&zero_bias, 256 * sizeof(float)));
    DBV_CUDA(cudaMemset(zero_bias, 0, 256 * sizeof(float)));
  }
  TcLayer T;
  memset(&T, 0, sizeof T);
  uint64_t dims[5] = {(uint64_t)K, 1, 1, 1, (uint64_t)M};
  uint64_t str[4] = {(uint64_t)K * 2, (uint64_t)K * 2, (uint64_t)K * 2, (uint64_t)K * 2};
  uint32_t box[5] = {(uint32_t)CBK, 1, 1, 1, 128};
  int r = encode_tmap(&T.tmA, a_dev, 5, dims, str, box, CBK * 2);
  if (r) return r;
  const int nkb = K / CBK;
  uint64_t bd[2] = {(uint64_t)CBK, (uint64_t)nkb * N};
  uint64_t bs[1] = {(uint64_t)CBK * 2};
  uint32_t bb[2] = {(uint32_t)CBK, (uint32_t)N};
  r = encode_tmap(&T.tmB, b_dev, 2, bd, bs, bb, CBK * 2);
  if (r) return r;
  for (int i = 0; i < nkb; ++i) {
    T.kb[i].dx = T.kb[i].dy = T.kb[i].plane = 0;
    T.kb[i].c_off = (int16_t)(i * CBK);
    T.kb[i].b_row = i * N;
  }
  T.n_cls = 1;
  T.cls[0] = TcClass{0, nkb, 0, 0, 1, 1};
  T.TW = T.TH = 1;
  T.TB = 128;
  T.SH = T.SW = 1;
  T.tiles_x = T.tiles_y = 1;
  T.n_tiles_n = 1;
  T.B = M;
  T.tiles_per_cls = (M + 127) / 128;
  T.total_tiles = T.tiles_per_cls;
  T.a_bytes = CBK * 2 * 128;
  T.b_bytes = N * CBK * 2;
  tc_stage_plan(T, CBK, N);
  T.dbg_shift_rows = shift;
  T.dbg_base_mode = base_mode;
  OutSpec& o = T.o;
  o.out = out_dev;
  o.mode = OUT_F32_NHWC;
  o.planes = 1;
  o.OH = o.OW = 1;
  o.Cout = o.Cpad = N;
  o.bias = zero_bias;
  return launch_tc_layer(T, CBK, N, kNumSMs, (cudaStream_t)stream);
}
