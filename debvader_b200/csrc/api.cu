// C-ABI of libdebvader_b200: context, weights, layer plans and the network entry points.
// See include/debvader_b200.h for the contract and the reference citations.
#include "kernels.h"
#include "host_stage.h"

#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <cstring>
#include <cmath>
#include <cuda_fp16.h>

namespace dbv {

std::atomic<long long> g_launches{0};

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

enum LayerKind { L_CONV = 0, L_CONVT = 1, L_DENSE = 2 };

// static description of one GEMM-shaped layer of the DC2 network
struct LayerDesc {
  const char* name;
  int kind, stride;
  int Hin, Cin, Hout, Cout;  // square images
  int enc;                   // 1: encoder model (layer_with_weights-0), 0: decoder model
  int wn, an, a2n;           // checkpoint indices of kernel/bias, alpha, second alpha (-1: none)
  int relu_head;             // decoder head
};

static const LayerDesc kLayers[] = {
    {"enc_conv1", L_CONV, 1, 59, 6, 59, 32, 1, 1, 2, -1, 0},
    {"enc_conv2", L_CONV, 2, 59, 32, 30, 32, 1, 3, 4, -1, 0},
    {"enc_conv3", L_CONV, 1, 30, 32, 30, 64, 1, 5, 6, -1, 0},
    {"enc_conv4", L_CONV, 2, 30, 64, 15, 64, 1, 7, 8, -1, 0},
    {"enc_conv5", L_CONV, 1, 15, 64, 15, 128, 1, 9, 10, -1, 0},
    {"enc_conv6", L_CONV, 2, 15, 128, 8, 128, 1, 11, 12, -1, 0},
    {"enc_conv7", L_CONV, 1, 8, 128, 8, 256, 1, 13, 14, -1, 0},
    {"enc_conv8", L_CONV, 2, 8, 256, 4, 256, 1, 15, 16, 17, 0},  // + Flatten PReLU (alpha2)
    {"enc_dense", L_DENSE, 1, 1, 4096, 1, 560, 1, 18, -1, -1, 0},
    {"dec_dense1", L_DENSE, 1, 1, 32, 1, 560, 0, 1, 2, -1, 0},
    {"dec_dense2", L_DENSE, 1, 1, 560, 1, 4096, 0, 3, 4, -1, 0},
    {"dec_convT1", L_CONVT, 2, 4, 256, 8, 256, 0, 5, 6, -1, 0},
    {"dec_convT2", L_CONVT, 1, 8, 256, 8, 256, 0, 7, 8, -1, 0},
    {"dec_convT3", L_CONVT, 2, 8, 256, 16, 128, 0, 9, 10, -1, 0},
    {"dec_convT4", L_CONVT, 1, 16, 128, 16, 128, 0, 11, 12, -1, 0},
    {"dec_convT5", L_CONVT, 2, 16, 128, 32, 64, 0, 13, 14, -1, 0},
    {"dec_convT6", L_CONVT, 1, 32, 64, 32, 64, 0, 15, 16, -1, 0},
    {"dec_convT7", L_CONVT, 2, 32, 64, 64, 32, 0, 17, 18, -1, 0},
    {"dec_convT8", L_CONVT, 1, 64, 32, 64, 32, 0, 19, 20, -1, 0},
    {"dec_head", L_CONV, 1, 64, 32, 64, 12, 0, 21, -1, -1, 1},
};
constexpr int kNumLayers = sizeof(kLayers) / sizeof(kLayers[0]);
// first layer of the decoder tail that reads a channel-group-planar input (19 = head only ... 17 = convT7, convT8, head; 20 = none)
static int cg8_first() {
  static const int v = dbv_env("DBV_CG8_FIRST") ? atoi(dbv_env("DBV_CG8_FIRST")) : 20;
  return v;
}
constexpr int PH_SMEM_BUDGET = 232448 - 1024 - 512 - 2048;
enum { I_CONV1 = 0, I_CONV8 = 7, I_ENC_DENSE = 8, I_DENSE1 = 9, I_DENSE2 = 10, I_T1 = 11, I_T6 = 16, I_HEAD = 19 };

static std::string wkey(int enc, int n, const char* nm) {
  char buf[128];
  snprintf(buf, sizeof buf, "layer_with_weights-%d/layer_with_weights-%d/%s", enc ? 0 : 1, n, nm);
  return buf;
}

static int same_pad_before(int n, int k, int s) {
  const int out = (n + s - 1) / s;
  const int total = std::max((out - 1) * s + k - n, 0);
  return total / 2;
}

// tensor-core configuration of a layer
struct TcGeom {
  int CBK, NT, TW, TH, TB;
  int tc;  // 0: runs on the SIMT kernel even in tensor-core modes
};
static const TcGeom kTcBase[kNumLayers] = {
    {16, 32, 59, 2, 1, 1},   // enc_conv1: BN + 8-channel bf16 packing by a SIMT pre-kernel, then the no-swizzle halo kernel
    {32, 32, 30, 4, 1, 1},   // enc_conv2
    {32, 64, 30, 4, 1, 1},   // enc_conv3
    {64, 64, 15, 8, 1, 1},   // enc_conv4
    {64, 128, 15, 8, 1, 1},  // enc_conv5
    {64, 128, 8, 8, 2, 1},   // enc_conv6
    {64, 256, 8, 8, 2, 1},   // enc_conv7
    {64, 256, 4, 4, 8, 1},   // enc_conv8
    {64, 112, 1, 1, 128, 1}, // enc_dense
    {0, 0, 0, 0, 0, 0},      // dec_dense1 (K=32): SIMT
    {64, 256, 1, 1, 128, 1}, // dec_dense2
    {64, 256, 4, 4, 8, 1},   // dec_convT1 (class space 4x4)
    {64, 256, 8, 8, 2, 1},   // dec_convT2
    {64, 128, 8, 8, 2, 1},   // dec_convT3 (class space 8x8)
    {64, 128, 16, 8, 1, 1},  // dec_convT4
    {64, 64, 16, 8, 1, 1},   // dec_convT5 (class space 16x16)
    {64, 64, 32, 4, 1, 1},   // dec_convT6
    {64, 32, 32, 4, 1, 1},   // dec_convT7 (class space 32x32)
    {32, 32, 64, 2, 1, 1},   // dec_convT8
    {32, 16, 64, 2, 1, 1},   // dec_head
};

// geometry actually used: layers reading a channel-group-planar input take K = 16 chunks (two 8-channel groups per MMA)
struct TcTable {
  TcGeom g[kNumLayers];
  TcTable() {
    for (int i = 0; i < kNumLayers; ++i) {
      g[i] = kTcBase[i];
      if (i >= cg8_first() && i <= 19) g[i].CBK = 16;
    }
  }
  const TcGeom& operator[](int i) const { return g[i]; }
};
static const TcTable kTc;
// DBV_PREC_FP32TC: the many-channel layers run on tc_conv_kernel with promoted partial sums; an epilogue thread then holds
// NT / 2 accumulators in registers, so the 256-channel layers are N-tiled by 128
static TcGeom tc_geom(int precision, int li) {
  TcGeom g = kTc[li];
  if (precision == DBV_PREC_FP32TC && g.NT == 256) g.NT = 128;
  return g;
}

struct LayerRt {
  // device weights
  float* w_gather = nullptr;       // SIMT gather form [taps][Cin][CoutP]
  __nv_bfloat16* w_packed = nullptr;  // tcgen05 packed blocks [nblk][Ntot][CBK]
  float* bias = nullptr;
  std::vector<float> bias_host;
  float* alpha = nullptr;
  float* alpha2 = nullptr;
  int CoutP = 0;
  // output activation buffer
  void* out = nullptr;
  size_t out_bytes_per_stamp = 0;
  OutSpec ospec{};
  // tcgen05 plan
  TcLayer tc{};
  bool has_tc = false;
  TcLayer tcp{};  // CTA-pair plan (many-channel layers)
  bool has_pair = false;
  PairHLayer ph{};  // CTA-pair plan with a per-chunk halo box (8x8 .. 16x16 stride-1 / transposed layers)
  bool has_pairh = false;
  // resident-halo plan (preferred when it exists)
  HaloLayer halo{};
  bool has_halo = false;
};

}  // namespace dbv

using namespace dbv;

struct dbv_ctx {
  int device = 0;
  int precision = DBV_PREC_FP32;
  long long chunk = 0;
  bool finalized = false;
  std::map<std::string, HostTensor> host_w;
  LayerRt rt[kNumLayers];
  float* bn_scale = nullptr;
  float* bn_shift = nullptr;
  float* dec_alpha0 = nullptr;
  // per-chunk scratch
  float* params = nullptr;  // [chunk][560]
  float* z = nullptr;       // [chunk][32]
  float* zp = nullptr;      // [chunk][32]
  LayerRt im2col;           // tensor-core modes: conv1's operand, BN'd input as bf16 [chunk][59][59][planes*8]
  __nv_bfloat16* conv1_wimg = nullptr;  // conv1 weights as a ready-made shared-memory image (no-swizzle core matrices)
  std::vector<void*> allocs;
  long long launches = 0;
  // host-buffer pipeline
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  void* stage_in[2] = {nullptr, nullptr};   // raw host dtype
  float* stage_x[2] = {nullptr, nullptr};
  float* stage_mean[2] = {nullptr, nullptr};
  float* stage_std[2] = {nullptr, nullptr};
  float* stage_z[2] = {nullptr, nullptr};
  float* stage_eps[2] = {nullptr, nullptr};
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  bool pipe_ready = false;
  // pageable host input: pinned staging (two slots of the largest piece) filled by a pool of host threads (host_stage.h)
  float* hstage[2] = {nullptr, nullptr};
  size_t hstage_elems = 0;
  dbv::HostStagePool* hpool = nullptr;
  // A ctx owns ONE set of activation buffers: whatever the device-pointer entry points last enqueued on a caller's stream
  // is marked by this event, and the host pipeline (which computes on its own stream) waits for it before its first piece
  cudaEvent_t ev_user = nullptr;
  int* ovf_host = nullptr;  // host-mapped flag the fp16-tail epilogues set when an activation saturates (DBV_PREC_MIXED)
  int* ovf_dev = nullptr;
  // profiling
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_names;
  int prof_n = 0;
  int prof_calls = 0;  // calls recorded since dbv_set_profiling(ctx, 1); the events of all of them are kept (up to kProfMaxEvents)
};

namespace dbv {

bool pdl_enabled() {
  static const bool on = dbv_env("DBV_PDL") ? atoi(dbv_env("DBV_PDL")) != 0 : false;
  return on;
}

static int dev_alloc(dbv_ctx* c, void** p, size_t bytes, bool zero) {
  DBV_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  c->allocs.push_back(*p);
  if (zero) DBV_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16));
  return DBV_OK;
}
template <typename T>
static int upload(dbv_ctx* c, T** dst, const std::vector<T>& v) {
  int r = dev_alloc(c, (void**)dst, v.size() * sizeof(T), false);
  if (r) return r;
  DBV_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DBV_OK;
}

static const HostTensor* find_w(dbv_ctx* c, const std::string& k) {
  auto it = c->host_w.find(k);
  return it == c->host_w.end() ? nullptr : &it->second;
}

static int check_shape(dbv_ctx* c, const std::string& k, std::initializer_list<int64_t> want, const HostTensor** out) {
  const HostTensor* t = find_w(c, k);
  if (!t) return fail(DBV_ERR_INVALID, "missing weight tensor '%s'", k.c_str());
  if (t->shape.size() != want.size() || !std::equal(want.begin(), want.end(), t->shape.begin())) {
    std::string got, exp;
    for (auto d : t->shape) got += std::to_string(d) + ",";
    for (auto d : want) exp += std::to_string(d) + ",";
    return fail(DBV_ERR_UNSUPPORTED, "weight '%s' has shape [%s], the DC2 architecture needs [%s]", k.c_str(), got.c_str(),
                exp.c_str());
  }
  *out = t;
  return DBV_OK;
}

// W(ky,kx,ci,co) accessor in *gather form* for a layer: value multiplying in[.., ci] for output co at tap (ky,kx)
// of the checkpoint tensor.  Conv2D kernels are HWIO, Conv2DTranspose kernels are HWOI, Dense is IO.
static inline float w_at(const LayerDesc& L, const HostTensor& W, int ky, int kx, int ci, int co) {
  if (L.kind == L_CONV) return W.data[(((size_t)ky * 3 + kx) * L.Cin + ci) * L.Cout + co];
  if (L.kind == L_CONVT) return W.data[(((size_t)ky * 3 + kx) * L.Cout + co) * L.Cin + ci];
  return W.data[(size_t)ci * L.Cout + co];
}

// tap table of a layer in gather form.  For each tap: which checkpoint (ky,kx) it uses and where its
// input pixel sits relative to the output pixel (tile space), plus parity plane for stride-2 Conv2D
// and output class for stride-2 Conv2DTranspose.
struct Tap {
  int ky, kx, dy, dx, plane, cls;
};
static std::vector<Tap> make_taps(const LayerDesc& L) {
  std::vector<Tap> t;
  if (L.kind == L_DENSE) {
    t.push_back({0, 0, 0, 0, 0, 0});
    return t;
  }
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      Tap a{ky, kx, 0, 0, 0, 0};
      if (L.kind == L_CONV && L.stride == 1) {
        a.dy = ky - 1;
        a.dx = kx - 1;
      } else if (L.kind == L_CONV) {  // stride 2: input row 2y + ky - pb lives in parity plane (ky-pb)&1 at row y + ((ky-pb)>>1)
        const int pb = same_pad_before(L.Hin, 3, 2);
        const int ry = ky - pb, rx = kx - pb;
        a.plane = (ry & 1) * 2 + (rx & 1);
        a.dy = ry >> 1;  // arithmetic shift: floor
        a.dx = rx >> 1;
      } else if (L.stride == 1) {  // Conv2DTranspose s1, pb=1: in row = y - ky + 1
        a.dy = 1 - ky;
        a.dx = 1 - kx;
      } else {  // Conv2DTranspose s2, pb=0: y = 2i + ky -> class py = ky&1, in row i - (ky-py)/2
        const int py = ky & 1, px = kx & 1;
        a.cls = py * 2 + px;
        a.dy = -((ky - py) / 2);
        a.dx = -((kx - px) / 2);
      }
      t.push_back(a);
    }
  return t;
}

static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// 16-bit storage format of a context's weights / activations: bf16, or fp16 for DBV_PREC_FP16X3
static inline uint16_t f2h16(int f16, float f) {
  if (!f16) return f2bf(f);
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static inline float h162f(int f16, uint16_t u) {
  if (!f16) return bf2f(u);
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}

// Layers that read a channel-group-planar (OUT_BF16_CG8) input: they exist only as resident-halo kernels.
static bool consumes_cg8(int li) { return li >= cg8_first() && li <= I_HEAD; }

// Storage of the activations ENTERING layer li (and the format of its weights) in the tensor-core modes.
// DBV_PREC_MIXED: the four large-image decoder layers read single-plane fp16 activations with fp16 hi/lo weights.
static bool mixed_tail(int precision, int li) { return precision == DBV_PREC_MIXED && li >= I_T6 && li <= I_HEAD; }
static int layer_f16(int precision, int li) { return (precision == DBV_PREC_FP16X3 || precision == DBV_PREC_FP32TC || mixed_tail(precision, li)) ? 1 : 0; }
static int layer_in_planes(int precision, int li) {
  if (precision == DBV_PREC_BF16) return 1;
  return mixed_tail(precision, li) ? 1 : 2;
}
static bool prec_x3(int precision) {
  return precision == DBV_PREC_BF16X3 || precision == DBV_PREC_FP16X3 || precision == DBV_PREC_MIXED || precision == DBV_PREC_FP32TC;
}

// describe how layer li's OUTPUT is stored in the tensor-core modes
static void tc_out_layout(int li, int precision, OutSpec* o) {
  const LayerDesc& L = kLayers[li];
  o->planes = li + 1 < kNumLayers ? layer_in_planes(precision, li + 1) : 1;
  o->f16 = li + 1 < kNumLayers ? layer_f16(precision, li + 1) : 0;
  o->OH = L.Hout;
  o->OW = L.Hout;
  o->Cout = L.Cout;
  o->Cpad = L.Cout;
  o->PH = o->PW = 0;
  o->mode = OUT_BF16_NHWC;
  const bool next_is_s2_conv = (li + 1 < kNumLayers) && kLayers[li + 1].kind == L_CONV && kLayers[li + 1].stride == 2;
  if (next_is_s2_conv) {
    o->mode = OUT_BF16_PARITY;
    o->PH = o->PW = (L.Hout + 1) / 2;
  }
  if (li + 1 < kNumLayers && consumes_cg8(li + 1)) o->mode = OUT_BF16_CG8;
  if (li == I_ENC_DENSE) { o->mode = OUT_F32_NHWC; o->planes = 1; }
  if (li == I_DENSE1) o->Cpad = 576;                    // K of dec_dense2 padded to 9 x 64
  if (li == I_DENSE2) { o->OH = o->OW = 4; o->Cout = o->Cpad = 256; }  // Reshape(4,4,256), model/model.py:119
  if (li == I_HEAD) { o->mode = OUT_HEAD; o->planes = 1; }
}

static size_t out_elems_per_stamp(const OutSpec& o) {
  if (o.mode == OUT_HEAD) return 0;
  if (o.mode == OUT_BF16_PARITY) return (size_t)4 * o.PH * o.PW * o.planes * o.Cpad;
  return (size_t)o.OH * o.OW * o.planes * o.Cpad;
}

static int build_tc_layer(dbv_ctx* c, int li) {
  const LayerDesc& L = kLayers[li];
  const TcGeom G = tc_geom(c->precision, li);
  LayerRt& R = c->rt[li];
  const bool x3 = prec_x3(c->precision);
  const HostTensor* W = find_w(c, wkey(L.enc, L.wn, "kernel"));
  // ---- input tensor (previous layer's output buffer) -------------------------------------------
  const LayerRt& P = (li == I_CONV1) ? c->im2col : c->rt[li - 1];
  const OutSpec& in = P.ospec;
  const int in_cpad = in.Cpad, in_planes = in.planes;
  // enc_dense reads conv8's (4,4,256) map: 16 "taps" at pixel offsets; everything else via make_taps
  std::vector<Tap> taps;
  if (li == I_ENC_DENSE) {
    for (int p = 0; p < 16; ++p) taps.push_back({p, 0, p / 4, p % 4, 0, 0});
  } else if (li == I_CONV1) {
    taps.push_back({0, 0, 0, 0, 0, 0});  // the 9 taps are already unrolled along K by the im2col pre-kernel
  } else {
    taps = make_taps(L);
  }
  const int cin_tap = (li == I_ENC_DENSE) ? 256 : (li == I_CONV1 ? 64 : L.Cin);  // channels per tap in the input tensor
  const int nchunk = (cin_tap + G.CBK - 1) / G.CBK;
  const int Ntot = ((L.Cout + G.NT - 1) / G.NT) * G.NT;
  const int parts_w = x3 ? 2 : 1;
  // ---- pack weights: block (tap, chunk, part) = [Ntot][CBK] bf16, K contiguous ------------------
  const size_t blk_elems = (size_t)Ntot * G.CBK;
  const size_t nblk = taps.size() * nchunk * parts_w;
  std::vector<uint16_t> packed(nblk * blk_elems, 0);
  for (size_t ti = 0; ti < taps.size(); ++ti)
    for (int ch = 0; ch < nchunk; ++ch)
      for (int n = 0; n < L.Cout; ++n)
        for (int k = 0; k < G.CBK; ++k) {
          const int ci = ch * G.CBK + k;
          if (ci >= cin_tap) continue;
          float w;
          if (li == I_ENC_DENSE) w = W->data[((size_t)taps[ti].ky * 256 + ci) * L.Cout + n];  // flat (h,w,c) index
          else if (li == I_CONV1) w = ci < 54 ? W->data[(size_t)ci * L.Cout + n] : 0.f;  // HWIO flattened: k = (ky*3+kx)*6 + band
          else w = w_at(L, *W, taps[ti].ky, taps[ti].kx, ci, n);
          const int f16 = layer_f16(c->precision, li);
          const uint16_t hi = f2h16(f16, w);
          const size_t b0 = ((ti * nchunk + ch) * parts_w) * blk_elems + (size_t)n * G.CBK + k;
          packed[b0] = hi;
          if (x3) packed[b0 + blk_elems] = f2h16(f16, w - h162f(f16, hi));
        }
  {
    std::vector<__nv_bfloat16> tmp(packed.size());
    memcpy(tmp.data(), packed.data(), packed.size() * 2);
    int r = upload(c, &R.w_packed, tmp);
    if (r) return r;
  }
  // ---- k-block table ------------------------------------------------------------------------------
  TcLayer& T = R.tc;
  memset(&T, 0, sizeof T);
  if ((x3 && in_planes < 2) || in.mode == OUT_BF16_CG8) {  // single-plane / channel-group-planar input: only the resident-halo kernel runs this layer;
    uint64_t bd[2] = {(uint64_t)G.CBK, (uint64_t)(nblk * Ntot)};  // it needs the packed weights' tensor map
    uint64_t bs[1] = {(uint64_t)G.CBK * 2};
    uint32_t bb[2] = {(uint32_t)G.CBK, (uint32_t)G.NT};
    return encode_tmap(&T.tmB, R.w_packed, 2, bd, bs, bb, G.CBK * 2);
  }
  const int ncls = (L.kind == L_CONVT && L.stride == 2) ? 4 : 1;
  T.n_cls = ncls;
  int nkb = 0;
  for (int cl = 0; cl < ncls; ++cl) {
    T.cls[cl].kb_begin = nkb;
    for (size_t ti = 0; ti < taps.size(); ++ti) {
      if (taps[ti].cls != cl) continue;
      for (int ch = 0; ch < nchunk; ++ch) {
        // one k-block per (tap, chunk); in the hi/lo split precisions the kernel loads the lo activation plane
        // (channel offset + lo_coff) and the lo weight block (row offset + lo_brow) into the same stage
        if (nkb >= TC_MAX_KB) return fail(DBV_ERR_UNSUPPORTED, "%s: k-block table overflow", L.name);
        TcKBlock& K = T.kb[nkb++];
        K.dx = (int16_t)taps[ti].dx;
        K.dy = (int16_t)taps[ti].dy;
        K.plane = (int16_t)taps[ti].plane;
        K.c_off = (int16_t)(ch * G.CBK);
        K.b_row = (int32_t)(((ti * nchunk + ch) * parts_w) * Ntot);
      }
    }
    T.cls[cl].nkb = nkb - T.cls[cl].kb_begin;
    const int py = cl >> 1, px = cl & 1;
    T.cls[cl].oy0 = ncls == 4 ? py : 0;
    T.cls[cl].ox0 = ncls == 4 ? px : 0;
    T.cls[cl].osy = T.cls[cl].osx = ncls == 4 ? 2 : 1;
  }
  // ---- geometry -----------------------------------------------------------------------------------
  T.TW = G.TW; T.TH = G.TH; T.TB = G.TB;
  const int space = (ncls == 4) ? L.Hin : L.Hout;  // class space of a s2 transposed conv = its input grid
  T.SH = T.SW = (L.kind == L_DENSE) ? 1 : space;
  T.tiles_x = (T.SW + T.TW - 1) / T.TW;
  T.tiles_y = (T.SH + T.TH - 1) / T.TH;
  T.n_tiles_n = Ntot / G.NT;
  T.nt_pixel_mode = (li == I_DENSE2) ? 256 / G.NT : 0;
  T.a_bytes = G.CBK * 2 * G.TW * G.TH * G.TB;
  T.b_bytes = G.NT * G.CBK * 2;
  T.x3 = x3 ? 1 : 0;
  T.ab_f16 = layer_f16(c->precision, li);
  T.lo_coff = in_cpad;
  T.lo_brow = Ntot;
  tc_stage_plan(T, G.CBK, G.NT);
  // ---- tensor maps ----------------------------------------------------------------------------------
  {
    // input viewed as (C, W, H, P, B)
    uint64_t dims[5], str[4];
    uint32_t box[5] = {(uint32_t)G.CBK, (uint32_t)G.TW, (uint32_t)G.TH, 1u, (uint32_t)G.TB};
    const uint64_t Ct = (uint64_t)in_planes * in_cpad;
    uint64_t Wd, Hd, Pd;
    if (in.mode == OUT_BF16_PARITY) { Wd = in.PW; Hd = in.PH; Pd = 4; }
    else { Wd = in.OW; Hd = in.OH; Pd = 1; }
    dims[0] = Ct; dims[1] = Wd; dims[2] = Hd; dims[3] = Pd; dims[4] = (uint64_t)c->chunk;
    str[0] = Ct * 2; str[1] = str[0] * Wd; str[2] = str[1] * Hd; str[3] = str[2] * Pd;
    int r = encode_tmap(&T.tmA, P.out, 5, dims, str, box, G.CBK * 2);
    if (r) return r;
    uint64_t bd[2] = {(uint64_t)G.CBK, (uint64_t)(nblk * Ntot)};
    uint64_t bs[1] = {(uint64_t)G.CBK * 2};
    uint32_t bb[2] = {(uint32_t)G.CBK, (uint32_t)G.NT};
    r = encode_tmap(&T.tmB, R.w_packed, 2, bd, bs, bb, G.CBK * 2);
    if (r) return r;
  }
  if (!tc_layer_supported(G.CBK, G.NT)) return fail(DBV_ERR_UNSUPPORTED, "%s: no kernel for CBK=%d NT=%d", L.name, G.CBK, G.NT);
  R.has_tc = true;
  if (c->precision == DBV_PREC_FP32TC) {
    // accumulate at most 128 values of K (x 3 products) per TMEM accumulator, then promote (tc_conv.cu); the layers whose
    // whole K is short (conv2/3, convT6-8, head: K <= 576) run on the resident-halo kernel with one chain
    if (tc_seg_supported(G.CBK, G.NT)) T.seg_kb = std::max(1, 128 / G.CBK);
    return DBV_OK;  // no CTA-pair plans in this precision
  }
  if (tc_pair_supported(G.CBK, G.NT) && Ntot % G.NT == 0 && !dbv_env("DBV_NO_PAIR")) {
    R.tcp = T;
    uint64_t bd[2] = {(uint64_t)G.CBK, (uint64_t)(nblk * Ntot)};
    uint64_t bs[1] = {(uint64_t)G.CBK * 2};
    uint32_t bb[2] = {(uint32_t)G.CBK, (uint32_t)(G.NT / 2)};
    int r = encode_tmap(&R.tcp.tmB, R.w_packed, 2, bd, bs, bb, G.CBK * 2);
    if (r) return r;
    tc_pair_stage_plan(R.tcp, G.CBK, G.NT);
    R.has_pair = true;
  }
  return DBV_OK;
}

// CTA-pair plan with a per-chunk halo box (tc_pairh.cu) for the stride-1 Conv2D / Conv2DTranspose layers (and stride-2
// transposed convs over their input grid) on 8..16-pixel maps with N = 128 / 256 in the hi/lo split precisions.
static int build_pairh_layer(dbv_ctx* c, int li) {
  const LayerDesc& L = kLayers[li];
  const TcGeom& G = kTc[li];
  LayerRt& R = c->rt[li];
  const bool x3 = prec_x3(c->precision);
  if (!x3 || c->precision == DBV_PREC_FP32TC || !R.has_tc || dbv_env("DBV_NO_PAIRH") || G.CBK != 64 || !tc_pairh_supported(G.NT) || L.kind == L_DENSE) return DBV_OK;
  if (L.kind == L_CONV && L.stride != 1) return DBV_OK;
  const OutSpec& in = c->rt[li - 1].ospec;
  const int ncls = (L.kind == L_CONVT && L.stride == 2) ? 4 : 1;
  const int space = (ncls == 4) ? L.Hin : L.Hout;  // tile space = the input grid
  if (in.mode != OUT_BF16_NHWC || in.planes != 2 || in.OW != space || space < 8 || space > 16 || L.Cin % 64 != 0 || ncls * G.NT > 512) return DBV_OK;
  PairHLayer& P = R.ph;
  memset(&P, 0, sizeof P);
  const int TB = space <= 8 ? 2 : 1;
  const int HB = space + 2;
  const std::vector<Tap> taps = make_taps(L);
  if (taps.size() > 16) return DBV_OK;
  const int nchunk = L.Cin / 64, parts = 2;
  const int Ntot = G.NT;  // conv layers are not N-tiled
  P.n_cls = ncls;
  P.cls_groups = (ncls * G.NT > 256) ? 2 : 1;  // keep the accumulators of a work item within 256 columns (double buffered)
  P.ab_f16 = layer_f16(c->precision, li);
  int nt = 0;
  std::vector<int> order;
  for (int cl = 0; cl < ncls; ++cl) {
    P.cls_begin[cl] = nt;
    for (size_t ti = 0; ti < taps.size(); ++ti) {
      if (taps[ti].cls != cl) continue;
      P.tap_aoff[nt] = ((taps[ti].dy + 1) * TB * 10 + taps[ti].dx + 1) * 128;
      P.tap_brow[nt] = (int)(ti * nchunk * parts) * Ntot;
      ++nt;
    }
    P.cls[cl] = R.tc.cls[cl];
  }
  P.cls_begin[ncls] = nt;
  P.nchunk = nchunk;
  P.chunk_brow = parts * Ntot;
  P.lo_brow = Ntot;
  P.lo_coff = in.Cpad;
  P.TB = TB;
  P.tiles_x = (space + 7) / 8;
  P.SW = P.SH = space;
  P.abox_tx = 64 * 2 * 10 * TB * HB;
  P.abox_bytes = ((P.abox_tx + 1023) / 1024) * 1024;
  const int bh = (G.NT / 2) * 128;
  P.a_stages = 2;
  P.b_stages = std::min(8, (PH_SMEM_BUDGET - P.a_stages * 2 * P.abox_bytes) / (2 * bh));
  if (P.b_stages < 2) return DBV_OK;
  P.tail_pad = 2048;  // the last 8-row group of a 15-row map reads a few rows past its box
  P.smem_bytes = 1024 + P.a_stages * 2 * P.abox_bytes + P.b_stages * 2 * bh + P.tail_pad + 512;
  const uint64_t Ct = (uint64_t)in.planes * in.Cpad;
  uint64_t dims[5] = {Ct, (uint64_t)in.OW, (uint64_t)c->chunk, (uint64_t)in.OH, 1};
  uint64_t str[4] = {Ct * 2, Ct * 2 * in.OW * in.OH, Ct * 2 * in.OW, Ct * 2 * in.OW * in.OH * (uint64_t)c->chunk};
  uint32_t box[5] = {64u, 10u, (uint32_t)TB, (uint32_t)HB, 1u};
  int r = encode_tmap(&P.tmA, c->rt[li - 1].out, 5, dims, str, box, 128);
  if (r) return r;
  {  // packed weights, box (64, NT/2): each CTA of the pair loads its half of a block
    const size_t nblk = taps.size() * nchunk * parts;
    uint64_t bd[2] = {64, (uint64_t)(nblk * Ntot)};
    uint64_t bs[1] = {128};
    uint32_t bb[2] = {64u, (uint32_t)(G.NT / 2)};
    if ((r = encode_tmap(&P.tmB, R.w_packed, 2, bd, bs, bb, 128))) return r;
  }
  R.has_pairh = true;
  return DBV_OK;
}

// Fill T with the plan for R output rows per band and nbuf halo buffers.  Returns 1 if the plan is valid
// (fits shared memory / TMEM / descriptor fields), 0 if not, < 0 on error.
static int halo_plan(dbv_ctx* c, int li, int bandR, int nbuf, int U, HaloLayer& T, bool pair = false) {
  const LayerDesc& L = kLayers[li];
  const TcGeom& G = kTc[li];
  LayerRt& R = c->rt[li];
  const LayerRt& P = (li == I_CONV1) ? c->im2col : c->rt[li - 1];
  const OutSpec& in = P.ospec;
  const bool x3 = prec_x3(c->precision);
  const bool c1 = (li == I_CONV1);  // no-swizzle mode: 16-byte pixel rows, K=16 = two adjacent pixels
  const bool cg8 = in.mode == OUT_BF16_CG8;       // channel-group-planar input: 16-byte pixel rows, K = 16 = two group planes
  const int ROWB = (c1 || cg8) ? 16 : G.CBK * 2;  // activation row pitch in shared memory
  const int ROWB_W = c1 ? 32 : G.CBK * 2;         // bytes of one weight row
  if (cg8 && (G.CBK != 16 || in.Cpad != L.Cin || L.Cin % 16 != 0)) return fail(DBV_ERR_STATE, "%s: channel-group-planar input needs K = 16 chunks", L.name);
  std::vector<Tap> taps = make_taps(L);
  if (c1) {  // per kernel row ky: pixel pairs (x-1, x) and (x+1, x+2); Tap.kx = pair index, dx = first pixel of the pair
    taps.clear();
    for (int ky = 0; ky < 3; ++ky)
      for (int pr = 0; pr < 2; ++pr) taps.push_back(Tap{ky, pr, ky - 1, 2 * pr - 1, 0, 0});
  }
  const int pad = 1;  // left zero column (slot 0 of every row is x = -1)
  int pad_top = 0, pad_bot = 0;  // halo rows above / below the band: what the taps actually reach (a stride-2 transposed conv
  for (const Tap& t : taps) {    // and a stride-2 conv look only one way)
    pad_top = std::max(pad_top, -t.dy);
    pad_bot = std::max(pad_bot, t.dy);
  }
  const int hrows = bandR + pad_top + pad_bot;
  const int nchunk = c1 ? 1 : (L.Cin + G.CBK - 1) / G.CBK;
  const int parts_w = x3 ? 2 : 1;
  const int ncls = (L.kind == L_CONVT && L.stride == 2) ? 4 : 1;
  // Row pitch of the halo tile: W + 1 — ONE shared zero column per row (slot 0 = x = -1): the right neighbour of x = W-1 is
  // the next row's slot 0, and the slot after the last row is zeroed slack.  conv1's pair trick (x+1, x+2) keeps W + 3.
  const int W = (ncls == 4) ? L.Hin : L.Hout, H = W, WP = c1 ? W + 3 : W + 1;
  // CTA-pair plan (tc_halo2.cu): each CTA of the pair keeps ONE part of every (tap, chunk) weight block — the leader hi, its peer
  // lo —, the MMAs have M = 256 and read the two halves of [B_hi | B_lo] from the two shared memories
  if (pair && (!x3 || in.planes != 1 || c1 || cg8 || (L.kind == L_CONVT && L.stride == 2) || in.mode == OUT_BF16_PARITY)) return 0;
  const int n_wblk = (int)taps.size() * nchunk * (pair ? 1 : parts_w);
  const int w_bytes = ((n_wblk * G.NT * ROWB_W + 1023) / 1024) * 1024;
  // stride-2 Conv2D: the input is stored as 4 parity planes (OUT_BF16_PARITY); every (plane, hi/lo, chunk) is its own region
  const int npar = (in.mode == OUT_BF16_PARITY) ? 4 : 1;
  const int n_regions = in.planes * nchunk * npar * (cg8 ? 2 : 1);  // CG8: two 8-channel group planes per K = 16 chunk
  if ((!cg8 && n_regions > 8) || bandR < 1 || bandR > H) return 0;
  const int ntiles = (bandR * WP + 127) / 128;
  const int CW = G.NT * (x3 ? 2 : 1);           // accumulator columns of one (class, tile): [hi-weight part | lo-weight part]
  // Stride-2 transposed conv: the four output-parity classes that share an INPUT SHIFT are concatenated along N — shift
  // (0,0) feeds all four classes (one MMA of N = 4*CW instead of four), (0,-1) and (-1,0) two each, (-1,-1) one: 4 MMAs per
  // k-step instead of 9, and the activation operand (whose shared-memory fetch bounds these few-channel layers) is read
  // 4 times instead of 9.  Column order of the classes inside a tile's accumulator: [2, 0, 1, 3], so that every shift's
  // class set is contiguous.
  const bool concat = ncls == 4 && 4 * CW <= 256 && !dbv_env("DBV_NO_CONCAT");
  // DBV_PREC_FP32TC: tcgen05 does not round its fp32 accumulation to nearest, so a chain of K = 288 .. 576 values x 3 products
  // costs 1.4e-6 .. 1.8e-6 of the output scale (tools/tc_accum_probe.cu).  The 3x3 taps of a stride-1 layer are therefore
  // spread over nseg accumulators side by side (chains of <= 96 .. 320 values of K) that the epilogue adds up in fp32
  // registers.  The stride-2 transposed convs have at most 4 taps per output class: one chain.
  int nseg = 1;
  if (c->precision == DBV_PREC_FP32TC && ncls == 1 && !c1) nseg = std::max(1, std::min(3, 256 / CW));
  const int DW = concat ? 4 * CW : nseg * CW;   // accumulator columns of a sub-unit
  const int nsub = concat ? ntiles : ncls * ntiles;  // sub-units per band (class-major when not concatenated)
  if (U * DW > 256 || (U != 1 && U != 2 && U != 4)) return 0;  // a unit (U sub-units) must fit 256 TMEM columns
  if (hrows > 256 || WP > 256 || ntiles > 32) return 0;
  int slot_cols = 32;
  while (slot_cols < U * DW) slot_cols *= 2;
  // >= one zeroed slack slot after the box; CG8: the regions of a buffer are ONE contiguous TMA box, the slack follows the buffer
  const long long region = cg8 ? (long long)hrows * WP * ROWB : (((long long)hrows * WP * ROWB + ROWB + 1023) / 1024) * 1024;
  const long long buf = cg8 ? ((n_regions * region + 16 + 1023) / 1024) * 1024 : n_regions * region;
  // layout: [halo ring][weights][barriers].  The last tile of a band over-reads < 131 garbage rows past its region: into the
  // next region, or (last region of the last buffer) into the weights — readable memory, results dropped by the epilogue.
  const int overread = 131 * ROWB;
  const int tail_pad = w_bytes >= overread ? 0 : ((overread - w_bytes + 1023) / 1024) * 1024;
  const long long smem = 1024 + w_bytes + (long long)nbuf * buf + tail_pad + 2048;  // + barriers
  if (smem > HALO_MAX_SMEM || nbuf < 1 || nbuf > 4) return 0;
  if (n_wblk > HALO_MAX_WBLK) return 0;
  memset(&T, 0, sizeof T);
  T.n_cls = ncls;
  for (int cl = 0; cl < ncls; ++cl) {
    T.cls[cl].oy0 = ncls == 4 ? (cl >> 1) : 0;
    T.cls[cl].ox0 = ncls == 4 ? (cl & 1) : 0;
    T.cls[cl].osy = T.cls[cl].osx = ncls == 4 ? 2 : 1;
  }
  const int f16 = layer_f16(c->precision, li);
  auto idesc_of = [&](int n) { return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24); };
  const long long MSTEP = 128ll * ROWB;  // bytes between consecutive 128-position tiles of a band
  const int ksteps = c1 ? 1 : G.CBK / 16;
  const long long wblk_bytes = (long long)G.NT * ROWB_W;
  // ---- resident weight layout -------------------------------------------------------------------------------------
  // plain: block j = packed block j = ((tap * nchunk + chunk) * parts + part).  concat: for every (shift, chunk) the
  // blocks of the classes using that shift, in column order, each as [hi, lo] — contiguous rows = one N-concatenated B operand.
  static const int col_order[4] = {2, 0, 1, 3};
  int ccol[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) ccol[col_order[i]] = i * CW;
  struct ShiftUse { int dy, dx; std::vector<std::pair<int, int>> cls_tap; };  // (class, tap index) in column order
  std::vector<ShiftUse> shifts;
  std::vector<long long> wblk_of;  // concat: smem block index of (shift s, chunk ch, k-th class of the shift, part) = wblk_of[s] + (ch * n_s + k) * parts
  if (concat) {
    for (int dy = 0; dy >= -1; --dy)
      for (int dx = 0; dx >= -1; --dx) {
        ShiftUse su{dy, dx, {}};
        for (int i = 0; i < 4; ++i)
          for (size_t ti = 0; ti < taps.size(); ++ti)
            if (taps[ti].cls == col_order[i] && taps[ti].dy == dy && taps[ti].dx == dx) su.cls_tap.push_back({col_order[i], (int)ti});
        shifts.push_back(su);
      }
    int j = 0;
    for (const ShiftUse& su : shifts) {
      wblk_of.push_back(j);
      for (int ch = 0; ch < nchunk; ++ch)
        for (const auto& ct : su.cls_tap)
          for (int part = 0; part < parts_w; ++part) T.w_src[j++] = (uint8_t)((ct.second * nchunk + ch) * parts_w + part);
    }
    if (j != n_wblk) return fail(DBV_ERR_STATE, "%s: concatenated weight layout has %d blocks, expected %d", L.name, j, n_wblk);
  } else if (pair) {
    for (int j = 0; j < n_wblk; ++j) T.w_src[j] = (uint8_t)(j * parts_w);  // + cluster rank in the kernel: hi block / lo block
  } else {
    for (int j = 0; j < n_wblk; ++j) T.w_src[j] = (uint8_t)j;
  }
  // ---- ops + items of one band --------------------------------------------------------------------------------------
  int nops = 0, nitems = 0, nunits = 0;
  const int NV = (G.NT % 32 != 0) ? 16 : 32;  // channels per epilogue item (must match tc_halo_kernel)
  const int NCHK = G.NT / NV;
  auto push_op = [&](long long a_off, long long b_off, int n, int dcol, bool acc) -> bool {
    if (nops >= HALO_MAX_OPS || (a_off >> 4) > 0x3fff || (b_off >> 4) > 0x3fff) return false;
    T.ops[nops].a = (uint32_t)(a_off >> 4);
    T.ops[nops].b = (uint32_t)(b_off >> 4);
    T.ops[nops].idesc = idesc_of(n);
    T.ops[nops].d = (uint32_t)dcol | (acc ? 0x10000u : 0u);
    ++nops;
    return true;
  };
  auto a_region = [&](int ch, int a_lo, int plane) -> long long { return cg8 ? (long long)(a_lo * nchunk + ch) * 2 : (long long)(a_lo * nchunk + ch) * npar + plane; };
  for (int s0 = 0; s0 < nsub; s0 += U) {
    if (nunits >= HALO_MAX_UNITS) return 0;
    const int s1 = std::min(s0 + U, nsub);
    for (int sidx = s0; sidx < s1; ++sidx) {
      const int dbase = (sidx - s0) * DW;
      if (!concat) {
        const int cl = sidx / ntiles, m = sidx % ntiles;
        bool first = true;
        int ntap_cl = 0, tap_no = 0, cur_seg = 0;
        for (size_t ti = 0; ti < taps.size(); ++ti) ntap_cl += taps[ti].cls == cl;
        for (size_t ti = 0; ti < taps.size(); ++ti) {
          if (taps[ti].cls != cl) continue;
          const int seg = tap_no++ * nseg / ntap_cl;  // accumulator of this tap
          if (seg != cur_seg) { cur_seg = seg; first = true; }
          const int dseg = dbase + seg * CW;
          // (A_hi x [B_hi | B_lo]) as ONE MMA of N = 2*NT (the hi and lo weight blocks are adjacent in shared memory)
          // + (A_lo x B_hi): A_hi is fetched once
          for (int ch = 0; ch < nchunk; ++ch)
            for (int pr = 0; pr < in.planes; ++pr)  // pr = 1: the lo activation plane (absent for single-plane inputs)
              for (int k = 0; k < ksteps; ++k) {
                const long long a_off = a_region(ch, pr, taps[ti].plane) * region + (long long)((taps[ti].dy + pad_top) * WP + taps[ti].dx + pad) * ROWB + 32 * k + m * MSTEP;
                const long long b_off = (long long)((ti * nchunk + ch) * (pair ? 1 : parts_w)) * wblk_bytes + 32 * k;
                if (!push_op(a_off, b_off, (x3 && pr == 0) ? 2 * G.NT : G.NT, dseg, !first)) return 0;
                first = false;
              }
        }
        for (int q = 0; q < NCHK; ++q) {
          if (nitems >= HALO_MAX_ITEMS) return 0;
          T.items[nitems++] = (uint32_t)(dbase + q * NV) | ((uint32_t)cl << 9) | ((uint32_t)m << 11) | ((uint32_t)q << 16);
        }
      } else {
        const int m = sidx;
        bool first = true;
        for (size_t si = 0; si < shifts.size(); ++si) {
          const ShiftUse& su = shifts[si];
          const int ns = (int)su.cls_tap.size();
          const int dcol0 = dbase + ccol[su.cls_tap[0].first];
          for (int ch = 0; ch < nchunk; ++ch)
            for (int pr = 0; pr < in.planes; ++pr)
              for (int k = 0; k < ksteps; ++k) {
                const long long a_off = a_region(ch, pr, 0) * region + (long long)((su.dy + pad_top) * WP + su.dx + pad) * ROWB + 32 * k + m * MSTEP;
                const long long blk0 = wblk_of[si] + (long long)ch * ns * parts_w;
                if (pr == 0) {  // A (hi) x [class blocks, each hi|lo]: one MMA of N = ns * CW
                  if (!push_op(a_off, blk0 * wblk_bytes + 32 * k, ns * CW, dcol0, !first)) return 0;
                  first = false;
                } else {        // A_lo x B_hi: the hi blocks are not adjacent, one MMA of N = NT per class
                  for (int kc = 0; kc < ns; ++kc)
                    if (!push_op(a_off, (blk0 + (long long)kc * parts_w) * wblk_bytes + 32 * k, G.NT, dbase + ccol[su.cls_tap[kc].first], true)) return 0;
                }
              }
        }
        for (int i = 0; i < 4; ++i)
          for (int q = 0; q < NCHK; ++q) {
            if (nitems >= HALO_MAX_ITEMS) return 0;
            const int cl = col_order[i];
            T.items[nitems++] = (uint32_t)(dbase + ccol[cl] + q * NV) | ((uint32_t)cl << 9) | ((uint32_t)m << 11) | ((uint32_t)q << 16);
          }
      }
    }
    T.unit_op_end[nunits] = (uint16_t)nops;
    T.unit_item_end[nunits] = (uint16_t)nitems;
    ++nunits;
  }
  T.n_units = nunits;
  T.slot_cols = slot_cols;
  T.ab_f16 = layer_f16(c->precision, li);
  for (int i = 0; i < 64; ++i) T.bias_c[i] = i < L.Cout ? R.bias_host[i] : 0.f;
  T.W = W; T.H = H; T.R = bandR; T.WP = WP; T.pad = pad; T.pad_top = pad_top;
  T.ntiles = ntiles;
  T.magic_wp = (uint32_t)(((1ull << 32) + WP - 1) / WP);
  T.n_regions = n_regions;
  for (int r = 0; r < n_regions; ++r) {  // r = ((plane_hi_lo * nchunk + chunk) * npar + parity plane)
    const int pc = r / npar;
    T.region_coff[r] = (pc / nchunk) * in.Cpad + (pc % nchunk) * G.CBK;
    T.region_c3[r] = r % npar;
  }
  T.w_img = c1 ? (const void*)c->conv1_wimg : nullptr;
  T.a_box_bytes = hrows * WP * ROWB;
  T.region_bytes = (int)region;
  T.buf_bytes = (int)buf;
  T.cg8 = cg8 ? 1 : 0;
  T.n_wblk = n_wblk;
  T.w_rows_per_blk = G.NT;  // conv layers are not N-tiled: Ntot == NT
  T.w_bytes = w_bytes;
  T.nbuf = nbuf;
  T.U = U;
  T.dbg_skip = dbv_env("DBV_HALO_SKIP") ? atoi(dbv_env("DBV_HALO_SKIP")) : 0;
  T.dbg_id = li;
  T.wide = x3 ? 1 : 0;
  T.nseg = nseg;
  T.seg_cols = CW;
  T.pair = pair ? 1 : 0;
  T.tail_pad = tail_pad;
  T.smem_bytes = (int)smem;
  T.bands_per_img = (H + bandR - 1) / bandR;
  const uint64_t Ct = (uint64_t)in.planes * in.Cpad;
  int r;
  if (cg8) {  // u64 elements: (2W, H, planes * Cin/8, B, 1); ONE box = whole rows of 16-byte pixels of every channel-group plane
    const uint64_t ng = Ct / 8;
    uint64_t dims[5] = {2ull * in.OW, (uint64_t)in.OH, ng, (uint64_t)c->chunk, 1};
    uint64_t str[4] = {16ull * in.OW, 16ull * in.OW * in.OH, 16ull * in.OW * in.OH * ng, 16ull * in.OW * in.OH * ng * c->chunk};
    uint32_t bx[5] = {(uint32_t)(2 * WP), (uint32_t)hrows, (uint32_t)ng, 1u, 1u};
    if (2 * WP > 256 || ng > 256) return 0;
    r = encode_tmap(&T.tmA, P.out, 5, dims, str, bx, 0, 8);
  } else {
    const uint64_t Wd = npar == 4 ? in.PW : in.OW, Hd = npar == 4 ? in.PH : in.OH;
    uint64_t dims[5] = {Ct, Wd, Hd, (uint64_t)npar, (uint64_t)c->chunk};
    uint64_t str[4] = {Ct * 2, Ct * 2 * Wd, Ct * 2 * Wd * Hd, Ct * 2 * Wd * Hd * npar};
    uint32_t box[5] = {(uint32_t)(c1 ? 8 : G.CBK), (uint32_t)WP, (uint32_t)hrows, 1u, 1u};
    r = encode_tmap(&T.tmA, P.out, 5, dims, str, box, c1 ? 0 : ROWB);
  }
  if (r) return r;
  if (!c1) T.tmB = R.tc.tmB;
  return 1;
}

// Resident-halo plan for the stride-1 Conv2D / Conv2DTranspose layers (and stride-2 transposed convs as 4
// classes over one halo) whose packed weights fit in shared memory next to the halo ring.  The band height
// and ring depth are AUTOTUNED on the device at finalize time: analytic models of the MMA / epilogue /
// barrier overlap mispredicted the best plan by up to 30 % ("halo plan autotune").
static int build_halo_layer(dbv_ctx* c, int li) {
  const LayerDesc& L = kLayers[li];
  const TcGeom& G = kTc[li];
  LayerRt& R = c->rt[li];
  if (dbv_env("DBV_NO_HALO") && li != I_CONV1) return DBV_OK;
  // DBV_PREC_FP32TC: convT6 has the longest chain of the resident-halo layers (K = 576) and, with N = 64, room for only two
  // partial accumulators per tile in TMEM (chains of 256 / 320): it runs on tc_conv_kernel with 128-value segments instead
  if (c->precision == DBV_PREC_FP32TC && li == I_T6) return DBV_OK;
  const bool must = li == I_CONV1 || mixed_tail(c->precision, li) || consumes_cg8(li);  // these layers have no other tensor-core kernel
  if ((!R.has_tc && !must) || L.kind == L_DENSE) return DBV_OK;
  if (!halo_layer_supported(G.CBK, G.NT)) return DBV_OK;
  const OutSpec& in = ((li == I_CONV1) ? c->im2col : c->rt[li - 1]).ospec;
  if (in.mode != OUT_BF16_NHWC && in.mode != OUT_BF16_PARITY && in.mode != OUT_BF16_CG8) return DBV_OK;
  if ((L.kind == L_CONV && L.stride == 2) != (in.mode == OUT_BF16_PARITY)) return DBV_OK;
  // resident weights must leave room for a useful halo band (else the streaming kernels are the better plan)
  if (in.mode == OUT_BF16_PARITY && (long long)9 * L.Cin * G.NT * 2 * (c->precision == DBV_PREC_BF16 ? 1 : 2) > 64 * 1024) return DBV_OK;
  const int ncls = (L.kind == L_CONVT && L.stride == 2) ? 4 : 1;
  const int H = (ncls == 4) ? L.Hin : L.Hout, WP = H + (li == I_CONV1 ? 3 : 1);
  // candidates: the tallest band for each tile count (R*WP just below a multiple of 128), both ring depths
  struct Cand { int r, nbuf, U; };
  std::vector<Cand> cand;
  for (int nt = 1; nt <= 16; ++nt) {
    int r = std::min(H, nt * 128 / WP);
    if (r < 1) continue;
    for (int nbuf = 3; nbuf >= 1; --nbuf)
      for (int U = 1; U <= 4; U *= 2) {
        if (U > ncls * nt && U > 1) continue;  // no band has that many sub-units
        bool dup = false;
        for (const Cand& q : cand) dup = dup || (q.r == r && q.nbuf == nbuf && q.U == U);
        if (!dup) cand.push_back({r, nbuf, U});
      }
  }
  // 16 stamps per SM: enough bands per CTA for a stable ranking AND a working set (input + output of the layer) well beyond
  // the 126 MB L2, as in a full 4096-stamp pass — with 592 stamps the large-image layers ran out of L2 and the ranking
  // did not carry over (round 2: convT7 flipped between two plans 10 % apart at full size)
  const long long Bt = std::min<long long>(c->chunk, 2368);
  struct TunerRes {  // released on every return path
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float *hm = nullptr, *hs = nullptr;
    ~TunerRes() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (hm) cudaFree(hm);
      if (hs) cudaFree(hs);
    }
  } tr;
  DBV_CUDA(cudaEventCreate(&tr.e0));
  DBV_CUDA(cudaEventCreate(&tr.e1));
  cudaEvent_t e0 = tr.e0, e1 = tr.e1;
  float best_ms = 1e30f;
  HaloLayer best{};
  bool found = false;
  OutSpec o = R.ospec;
  if (li == I_HEAD) {  // the head writes the caller's buffers: give the tuner scratch ones
    DBV_CUDA(cudaMalloc(&tr.hm, (size_t)Bt * STAMP_ELTS * 4));
    DBV_CUDA(cudaMalloc(&tr.hs, (size_t)Bt * STAMP_ELTS * 4));
    o.out = tr.hm;
    o.out2 = tr.hs;
  }
  // the stride-1 layers of the fp16 tail of DBV_PREC_MIXED also have a CTA-pair form (tc_halo2.cu): both are timed
  const bool try_pair = mixed_tail(c->precision, li) && ncls == 1 && halo_pair_supported(G.CBK, G.NT) && !dbv_env("DBV_NO_HALO_PAIR");
  for (int pass = 0; pass < (try_pair ? 2 : 1); ++pass)
  for (const Cand& cd : cand) {
    const int r = cd.r, nbuf = cd.nbuf;
    const bool pair = pass == 1;
    HaloLayer T;
    int ok = halo_plan(c, li, r, nbuf, cd.U, T, pair);
    if (ok < 0) return ok;
    if (!ok) continue;
    T.B = Bt;
    T.o = o;
    T.total_bands = (pair ? (Bt + 1) / 2 : Bt) * T.bands_per_img;
    float ms = 0.f;
    for (int it = 0; it < 4; ++it) {  // 1 warm-up + best of 3
      DBV_CUDA(cudaEventRecord(e0, 0));
      int rr = pair ? launch_halo_pair_layer(T, G.CBK, G.NT, kNumSMs, 0) : launch_halo_layer(T, G.CBK, G.NT, kNumSMs, 0);
      if (rr) return rr;
      DBV_CUDA(cudaEventRecord(e1, 0));
      DBV_CUDA(cudaEventSynchronize(e1));
      float t;
      DBV_CUDA(cudaEventElapsedTime(&t, e0, e1));
      ms = (it == 1) ? t : (it >= 2 ? std::min(ms, t) : ms);
    }
    if (getenv("DBV_VERBOSE")) fprintf(stderr, "[dbv] %s: halo candidate%s R=%d nbuf=%d U=%d ntiles=%d smem=%d -> %.3f ms / %lld stamps\n", L.name, pair ? " (CTA pair)" : "", r, nbuf, cd.U, T.ntiles, T.smem_bytes, ms, Bt);
    if (ms < best_ms) { best_ms = ms; best = T; found = true; }
  }
  if (!found && must) return fail(DBV_ERR_STATE, "%s: no valid halo plan (there is no other tensor-core kernel for this layer in this precision)", L.name);
  if (!found) return DBV_OK;
  R.halo = best;
  R.has_halo = true;
  if (getenv("DBV_VERBOSE"))
    fprintf(stderr, "[dbv] %s: halo plan%s R=%d nbuf=%d U=%d ntiles=%d regions=%d smem=%d (%.3f ms / %lld stamps)\n", L.name, best.pair ? " (CTA pair)" : "", best.R, best.nbuf, best.U,
            best.ntiles, best.n_regions, best.smem_bytes, best_ms, Bt);
  return DBV_OK;
}

static void prof_mark(dbv_ctx* c, const char* name, cudaStream_t st);

static int run_layer(dbv_ctx* c, int li, const void* input_f32, long long B, float* head_mean, float* head_std, float* params_out,
                     cudaStream_t st) {
  const LayerDesc& L = kLayers[li];
  LayerRt& R = c->rt[li];
  OutSpec o = R.ospec;
  if (li == I_HEAD) { o.out = head_mean; o.out2 = head_std; }
  if (li == I_ENC_DENSE && params_out) o.out = params_out;
  if (li == I_CONV1 && R.has_halo) {
    int r = launch_bn_pack8((const float*)input_f32, c->bn_scale, c->bn_shift, B, c->im2col.ospec, st);
    if (r) return r;
    prof_mark(c, "enc_bn_pack", st);
  }
  if (R.has_halo) {
    HaloLayer T = R.halo;
    T.B = B;
    T.o = o;
    T.total_bands = (T.pair ? (B + 1) / 2 : B) * T.bands_per_img;
    if (T.pair) return launch_halo_pair_layer(T, kTc[li].CBK, kTc[li].NT, kNumSMs, st);
    return launch_halo_layer(T, kTc[li].CBK, kTc[li].NT, kNumSMs, st);
  }
  if (R.has_pairh) {
    PairHLayer T = R.ph;
    T.B = B;
    T.o = o;
    const long long mt = ((B + T.TB - 1) / T.TB) * T.tiles_x;
    T.pair_items = (mt + 1) / 2;
    return launch_tc_pairh(T, kTc[li].NT, kNumSMs, st);
  }
  if (R.has_pair) {
    TcLayer T = R.tcp;
    T.B = B;
    T.o = o;
    const long long mt = ((B + T.TB - 1) / T.TB) * T.tiles_x * T.tiles_y;
    T.pair_items = (long long)T.n_cls * ((mt + 1) / 2) * T.n_tiles_n;
    return launch_tc_pair(T, kTc[li].CBK, kTc[li].NT, kNumSMs, st);
  }
  if (R.has_tc) {
    TcLayer T = R.tc;
    T.B = B;
    T.o = o;
    const long long btiles = (B + T.TB - 1) / T.TB;
    T.tiles_per_cls = btiles * T.tiles_x * T.tiles_y * T.n_tiles_n;
    T.total_tiles = T.tiles_per_cls * T.n_cls;
    const TcGeom G = tc_geom(c->precision, li);
    return launch_tc_layer(T, G.CBK, G.NT, kNumSMs, st);
  }
  SimtConv p{};
  p.in = (const float*)input_f32;
  p.w = R.w_gather;
  p.B = B;
  p.Hin = p.Win = L.Hin;
  p.Cin = L.Cin;
  p.Hout = p.Wout = L.Hout;
  p.CoutP = R.CoutP;
  p.ksz = L.kind == L_DENSE ? 1 : 3;
  p.mode = (L.kind == L_CONVT && L.stride == 2) ? 1 : 0;
  p.stride = (L.kind == L_CONV) ? L.stride : 1;
  p.pb = L.kind == L_DENSE ? 0 : (L.kind == L_CONV ? same_pad_before(L.Hin, 3, L.stride) : 1);
  if (li == I_CONV1) { p.in_scale = c->bn_scale; p.in_shift = c->bn_shift; }
  p.o = o;
  return launch_simt_conv(p, st);
}

constexpr int kProfMaxEvents = 16384;
static void prof_mark(dbv_ctx* c, const char* name, cudaStream_t st) {
  if (!c->profiling || c->prof_n >= kProfMaxEvents) return;
  if ((int)c->prof_ev.size() <= c->prof_n) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->prof_ev.push_back(e);
    c->prof_names.push_back("");
  }
  c->prof_names[c->prof_n] = name;
  cudaEventRecord(c->prof_ev[c->prof_n], st);
  c->prof_n++;
}

// one chunk (B <= ctx->chunk) through the requested stages
static int run_chunk(dbv_ctx* c, const float* x, long long B, const float* eps, uint64_t seed, int sample, long long first_stamp,
                     float* params_out, float* z_io, float* loc_out, float* zstd_out, float* mean, float* stddev, bool do_enc,
                     bool do_lat, bool do_dec, cudaStream_t st) {
  const bool fp32 = c->precision == DBV_PREC_FP32;
  int r;
  if (c->profiling) prof_mark(c, "start", st);
  float* params = c->params;
  if (do_enc) {
    const void* in = x;
    for (int li = I_CONV1; li <= I_ENC_DENSE; ++li) {
      float* pout = (li == I_ENC_DENSE) ? (params_out && !do_lat ? params_out : params) : nullptr;
      r = run_layer(c, li, in, B, nullptr, nullptr, pout, st);
      if (r) return r;
      prof_mark(c, kLayers[li].name, st);
      in = fp32 ? c->rt[li].out : nullptr;  // tensor-core layers find their input through the tensor map
    }
    if (params_out && do_lat) DBV_CUDA(cudaMemcpyAsync(params_out, params, (size_t)B * NPAR * 4, cudaMemcpyDeviceToDevice, st));
  }
  float* z = c->z;
  if (do_lat) {
    const float* pin = do_enc ? params : params_out;  // dbv_latent passes the caller's params through params_out
    r = launch_latent(pin, eps, seed, sample, first_stamp, B, z_io ? z_io : z, loc_out, zstd_out, c->zp, c->dec_alpha0, st);
    if (r) return r;
    prof_mark(c, "latent", st);
  } else if (do_dec) {
    r = launch_prelu_vec(z_io, c->dec_alpha0, B * LAT, LAT, c->zp, st);
    if (r) return r;
  }
  if (do_dec) {
    const void* in = c->zp;
    for (int li = I_DENSE1; li <= I_HEAD; ++li) {
      // in tensor-core modes dense1 runs on the SIMT kernel from zp, the rest read through tensor maps
      r = run_layer(c, li, in, B, mean, stddev, nullptr, st);
      if (r) return r;
      prof_mark(c, kLayers[li].name, st);
      in = fp32 ? c->rt[li].out : nullptr;
    }
  }
  return DBV_OK;
}

static int check_ready(dbv_ctx* c, const char* fn) {
  if (!c) return fail(DBV_ERR_INVALID, "%s: null ctx", fn);
  if (!c->finalized) return fail(DBV_ERR_STATE, "%s: weights not finalized (call dbv_finalize_weights)", fn);
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return fail(DBV_ERR_CUDA, "%s: cudaSetDevice(%d): %s", fn, c->device, cudaGetErrorString(e));
  return DBV_OK;
}

}  // namespace dbv

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" int dbv_abi_version(void) { return DBV_ABI_VERSION; }
extern "C" const char* dbv_last_error(void) { return last_error().c_str(); }
extern "C" int64_t dbv_global_launch_count(void) { return g_launches.load(); }
extern "C" int64_t dbv_launch_count(const dbv_ctx* c) { return c ? c->launches : 0; }

extern "C" int dbv_create(dbv_ctx** out, int device, int precision, int64_t chunk) {
  DBV_REQUIRE(out, "dbv_create: null out");
  DBV_REQUIRE(precision >= DBV_PREC_FP32 && precision <= DBV_PREC_FP32TC, "dbv_create: bad precision %d", precision);
  DBV_REQUIRE(chunk >= 0 && chunk <= (1 << 20), "dbv_create: bad chunk %lld", (long long)chunk);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(DBV_ERR_CUDA, "dbv_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  DBV_REQUIRE(device >= 0 && device < ndev, "dbv_create: device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  DBV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(DBV_ERR_UNSUPPORTED, "dbv_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  DBV_CUDA(cudaSetDevice(device));
  dbv_ctx* c = new dbv_ctx();
  c->device = device;
  c->precision = precision;
  c->chunk = chunk > 0 ? chunk : 4096;  // stamps per pass of the layer sequence: large enough that a launch of the small-image
                                        // layers (one pass over the layer's weights per SM, ~50 us) is amortised
  *out = c;
  return DBV_OK;
}

extern "C" int dbv_destroy(dbv_ctx* c) {
  if (!c) return DBV_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->allocs) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
  }
  for (auto e : c->prof_ev) cudaEventDestroy(e);
  if (c->ev_user) cudaEventDestroy(c->ev_user);
  if (c->ovf_host) cudaFreeHost(c->ovf_host);
  for (int i = 0; i < 2; ++i)
    if (c->hstage[i]) cudaFreeHost(c->hstage[i]);
  delete c->hpool;
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_comp) cudaStreamDestroy(c->s_comp);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
  return DBV_OK;
}

extern "C" int dbv_set_weights(dbv_ctx* c, const char* key, const float* host, const int64_t* shape, int ndim) {
  DBV_REQUIRE(c && key && host && shape, "dbv_set_weights: null argument");
  DBV_REQUIRE(ndim >= 1 && ndim <= 4, "dbv_set_weights: bad ndim %d for '%s'", ndim, key);
  if (c->finalized) return fail(DBV_ERR_STATE, "dbv_set_weights: ctx already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DBV_REQUIRE(shape[i] > 0, "dbv_set_weights: bad shape for '%s'", key);
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(host, host + n);
  std::string k(key);
  const std::string suffix = "/.ATTRIBUTES/VARIABLE_VALUE";
  if (k.size() > suffix.size() && k.compare(k.size() - suffix.size(), suffix.size(), suffix) == 0) k.resize(k.size() - suffix.size());
  c->host_w[k] = std::move(t);
  return DBV_OK;
}

extern "C" int dbv_finalize_weights(dbv_ctx* c) {
  DBV_REQUIRE(c, "dbv_finalize_weights: null ctx");
  if (c->finalized) return fail(DBV_ERR_STATE, "dbv_finalize_weights: already finalized");
  DBV_CUDA(cudaSetDevice(c->device));
  const bool fp32 = c->precision == DBV_PREC_FP32;
  const int planes = prec_x3(c->precision) ? 2 : 1;  // conv1's operand (and every layer outside the mixed tail)
  const int f16 = (c->precision == DBV_PREC_FP16X3 || c->precision == DBV_PREC_FP32TC) ? 1 : 0;
  int r;
  // ---- BatchNorm (model/model.py:79; Keras eps 1e-3) folded to scale/shift -----------------------
  {
    const HostTensor *g, *b, *m, *v;
    if ((r = check_shape(c, wkey(1, 0, "gamma"), {6}, &g))) return r;
    if ((r = check_shape(c, wkey(1, 0, "beta"), {6}, &b))) return r;
    if ((r = check_shape(c, wkey(1, 0, "moving_mean"), {6}, &m))) return r;
    if ((r = check_shape(c, wkey(1, 0, "moving_variance"), {6}, &v))) return r;
    std::vector<float> sc(8, 0.f), sh(8, 0.f);
    for (int i = 0; i < 6; ++i) {
      sc[i] = g->data[i] / sqrtf(v->data[i] + 1e-3f);
      sh[i] = b->data[i] - m->data[i] * sc[i];
    }
    if ((r = upload(c, &c->bn_scale, sc))) return r;
    if ((r = upload(c, &c->bn_shift, sh))) return r;
    const HostTensor* a0;
    if ((r = check_shape(c, wkey(0, 0, "alpha"), {32}, &a0))) return r;
    if ((r = upload(c, &c->dec_alpha0, a0->data))) return r;
  }
  if (c->precision == DBV_PREC_MIXED) {
    DBV_CUDA(cudaHostAlloc((void**)&c->ovf_host, sizeof(int), cudaHostAllocMapped));
    *c->ovf_host = 0;
    DBV_CUDA(cudaHostGetDevicePointer((void**)&c->ovf_dev, c->ovf_host, 0));
  }
  // ---- per-layer weights ------------------------------------------------------------------------
  for (int li = 0; li < kNumLayers; ++li) {
    const LayerDesc& L = kLayers[li];
    LayerRt& R = c->rt[li];
    const HostTensor *W, *Bv, *A = nullptr, *A2 = nullptr;
    if (L.kind == L_CONV) r = check_shape(c, wkey(L.enc, L.wn, "kernel"), {3, 3, L.Cin, L.Cout}, &W);
    else if (L.kind == L_CONVT) r = check_shape(c, wkey(L.enc, L.wn, "kernel"), {3, 3, L.Cout, L.Cin}, &W);
    else r = check_shape(c, wkey(L.enc, L.wn, "kernel"), {L.Cin, L.Cout}, &W);
    if (r) return r;
    if ((r = check_shape(c, wkey(L.enc, L.wn, "bias"), {L.Cout}, &Bv))) return r;
    if (L.an >= 0) {
      if (L.kind == L_DENSE) r = check_shape(c, wkey(L.enc, L.an, "alpha"), {L.Cout}, &A);
      else r = check_shape(c, wkey(L.enc, L.an, "alpha"), {L.Hout, L.Hout, L.Cout}, &A);
      if (r) return r;
    }
    if (L.a2n >= 0 && (r = check_shape(c, wkey(L.enc, L.a2n, "alpha"), {(int64_t)L.Hout * L.Hout * L.Cout}, &A2))) return r;
    if ((r = upload(c, &R.bias, Bv->data))) return r;
    R.bias_host = Bv->data;
    R.CoutP = (L.Cout + 3) & ~3;
    const bool simt = fp32 || !kTc[li].tc;
    if (simt) {
      const int ksz = L.kind == L_DENSE ? 1 : 3;
      std::vector<float> g((size_t)ksz * ksz * L.Cin * R.CoutP, 0.f);
      for (int ky = 0; ky < ksz; ++ky)
        for (int kx = 0; kx < ksz; ++kx) {
          // stride-1 transposed conv = conv with the kernel flipped (in row = y + (2-ky) - 1)
          const bool flip = (L.kind == L_CONVT && L.stride == 1);
          const int sy = flip ? 2 - ky : ky, sx = flip ? 2 - kx : kx;
          for (int ci = 0; ci < L.Cin; ++ci)
            for (int co = 0; co < L.Cout; ++co)
              g[(((size_t)ky * ksz + kx) * L.Cin + ci) * R.CoutP + co] = w_at(L, *W, sy, sx, ci, co);
        }
      if ((r = upload(c, &R.w_gather, g))) return r;
    }
    // output buffer + spec
    OutSpec& o = R.ospec;
    memset(&o, 0, sizeof o);
    if (fp32) {
      o.mode = (li == I_HEAD) ? OUT_HEAD : OUT_F32_NHWC;
      o.planes = 1;
      o.OH = o.OW = L.Hout;
      o.Cout = o.Cpad = L.Cout;
    } else {
      tc_out_layout(li, c->precision, &o);
    }
    // PReLU slopes: checkpoint layout (h,w,c) of the map THIS OutSpec describes -> [c/4][h*w][4] (epilogue.cuh:alpha_index)
    auto regroup = [&](const HostTensor* t) {
      const long long npix = (long long)o.OH * o.OW;
      std::vector<float> g(t->data.size());
      for (long long pix = 0; pix < npix; ++pix)
        for (int ch = 0; ch < o.Cout; ++ch) g[alpha_index(npix, pix, ch)] = t->data[pix * o.Cout + ch];
      return g;
    };
    if ((A || A2) && (o.Cout % 4 != 0 || (size_t)o.OH * o.OW * o.Cout != (A ? A : A2)->data.size()))
      return fail(DBV_ERR_UNSUPPORTED, "%s: PReLU slopes do not match the output map", L.name);
    if (A && (r = upload(c, &R.alpha, regroup(A)))) return r;
    if (A2 && (r = upload(c, &R.alpha2, regroup(A2)))) return r;
    o.bias = R.bias;
    o.alpha = R.alpha;
    o.alpha2 = R.alpha2;
    o.alpha_le1 = 1;  // every slope <= 1 (NaN counts as "no"): the epilogues may then use prelu(v) = max(v, a v)
    for (const HostTensor* t : {A, A2})
      if (t)
        for (float a : t->data)
          if (!(a <= 1.0f)) o.alpha_le1 = 0;
    o.relu = L.relu_head;
    o.ovf = (!fp32 && o.mode == OUT_BF16_NHWC && o.planes == 1 && o.f16 == 1 && c->precision == DBV_PREC_MIXED) ? c->ovf_dev : nullptr;
    const size_t el = out_elems_per_stamp(o);
    const size_t esz = (o.mode == OUT_F32_NHWC) ? 4 : 2;
    R.out_bytes_per_stamp = el * esz;
    if (li == I_ENC_DENSE) {
      if ((r = dev_alloc(c, (void**)&c->params, (size_t)c->chunk * NPAR * 4, true))) return r;
      R.out = c->params;
    } else if (el) {
      if ((r = dev_alloc(c, &R.out, (size_t)c->chunk * R.out_bytes_per_stamp, true))) return r;
    }
    o.out = R.out;
  }
  if ((r = dev_alloc(c, (void**)&c->z, (size_t)c->chunk * LAT * 4, true))) return r;
  if ((r = dev_alloc(c, (void**)&c->zp, (size_t)c->chunk * LAT * 4, true))) return r;
  if (!fp32) {
    OutSpec& o = c->im2col.ospec;
    memset(&o, 0, sizeof o);
    o.mode = OUT_BF16_NHWC;
    o.planes = planes;
    o.f16 = f16;
    o.OH = o.OW = S_;
    o.Cout = o.Cpad = 8;
    if ((r = dev_alloc(c, &c->im2col.out, (size_t)c->chunk * S_ * S_ * planes * 8 * 2, true))) return r;
    o.out = c->im2col.out;
    // conv1 weights as the shared-memory image of the no-swizzle K-major B operand: block (ky, pair)[hi|lo] of
    // 32 rows x 16 k; element (n, k) at (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2 bytes; k<8: kx = 2*pair,
    // k>=8: kx = 2*pair+1 (zero when kx > 2), band = k%8 (zero for the two padding channels)
    const HostTensor* W1 = find_w(c, wkey(1, 1, "kernel"));
    const int parts = planes;
    std::vector<uint16_t> img((size_t)6 * parts * 32 * 16, 0);
    for (int ky = 0; ky < 3; ++ky)
      for (int pr = 0; pr < 2; ++pr)
        for (int n = 0; n < 32; ++n)
          for (int k = 0; k < 16; ++k) {
            const int kx = 2 * pr + (k >> 3), ch = k & 7;
            const float w = (kx <= 2 && ch < 6) ? W1->data[(((size_t)ky * 3 + kx) * 6 + ch) * 32 + n] : 0.f;
            const uint16_t hi = f2h16(f16, w);
            const size_t blk = (size_t)(ky * 2 + pr) * parts;
            const size_t e = (size_t)(n / 8) * 128 + (size_t)(k / 8) * 64 + (size_t)(n % 8) * 8 + (k % 8);  // in bf16 elements
            img[blk * 512 + e] = hi;
            if (parts == 2) img[(blk + 1) * 512 + e] = f2h16(f16, w - h162f(f16, hi));
          }
    std::vector<__nv_bfloat16> tmp(img.size());
    memcpy(tmp.data(), img.data(), img.size() * 2);
    if ((r = upload(c, &c->conv1_wimg, tmp))) return r;
  }
  // ---- tensor-core plans ------------------------------------------------------------------------
  if (!fp32)
    for (int li = 0; li < kNumLayers; ++li)
      if (kTc[li].tc && ((li != I_CONV1 && (r = build_tc_layer(c, li))) || (li != I_CONV1 && (r = build_pairh_layer(c, li))) || (r = build_halo_layer(c, li)))) return r;
  DBV_CUDA(cudaDeviceSynchronize());
  c->finalized = true;
  c->host_w.clear();
  return DBV_OK;
}

static int chunked(dbv_ctx* c, const char* fn, const float* x, int64_t B, const float* eps, uint64_t seed, int sample,
                   int64_t first_stamp, float* params, float* z, float* loc, float* zstd, float* mean, float* stddev, bool enc,
                   bool lat, bool dec, void* stream) {
  int r = check_ready(c, fn);
  if (r) return r;
  DBV_REQUIRE(B >= 0, "%s: negative B", fn);
  cudaStream_t st = (cudaStream_t)stream;
  const long long before = g_launches.load();
  if (c->profiling) c->prof_calls++;
  for (int64_t b0 = 0; b0 < B; b0 += c->chunk) {
    const long long nb = std::min<long long>(c->chunk, B - b0);
    r = run_chunk(c, x ? x + b0 * STAMP_ELTS : nullptr, nb, eps ? eps + b0 * LAT : nullptr, seed, sample, first_stamp + b0,
                  params ? params + b0 * NPAR : nullptr, z ? z + b0 * LAT : nullptr, loc ? loc + b0 * LAT : nullptr,
                  zstd ? zstd + b0 * LAT : nullptr, mean ? mean + b0 * STAMP_ELTS : nullptr,
                  stddev ? stddev + b0 * STAMP_ELTS : nullptr, enc, lat, dec, st);
    if (r) return r;
  }
  if (B > 0) {
    if (!c->ev_user) DBV_CUDA(cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming));
    DBV_CUDA(cudaEventRecord(c->ev_user, st));
  }
  c->launches += g_launches.load() - before;
  return DBV_OK;
}

extern "C" int dbv_encode(dbv_ctx* c, const float* x, int64_t B, float* params, void* stream) {
  DBV_REQUIRE(B == 0 || (x && params), "dbv_encode: null buffer");
  return chunked(c, "dbv_encode", x, B, nullptr, 0, 0, 0, params, nullptr, nullptr, nullptr, nullptr, nullptr, true, false, false, stream);
}

extern "C" int dbv_latent(dbv_ctx* c, const float* params, const float* eps, uint64_t seed, int sample, int64_t first_stamp,
                          int64_t B, float* z, float* loc, float* zstd, void* stream) {
  DBV_REQUIRE(B == 0 || (params && z), "dbv_latent: null buffer");
  return chunked(c, "dbv_latent", nullptr, B, eps, seed, sample, first_stamp, const_cast<float*>(params), z, loc, zstd, nullptr,
                 nullptr, false, true, false, stream);
}

extern "C" int dbv_decode(dbv_ctx* c, const float* z, int64_t B, float* mean, float* stddev, void* stream) {
  DBV_REQUIRE(B == 0 || (z && mean), "dbv_decode: null buffer");
  return chunked(c, "dbv_decode", nullptr, B, nullptr, 0, 0, 0, nullptr, const_cast<float*>(z), nullptr, nullptr, mean, stddev,
                 false, false, true, stream);
}

extern "C" int dbv_deblend(dbv_ctx* c, const float* x, int64_t B, const float* eps, uint64_t seed, int sample, float* mean,
                           float* stddev, float* z, void* stream) {
  DBV_REQUIRE(B == 0 || (x && mean), "dbv_deblend: null buffer");
  return chunked(c, "dbv_deblend", x, B, eps, seed, sample, 0, nullptr, z, nullptr, nullptr, mean, stddev, true, true, true, stream);
}

// ---- host-buffer pipeline ----------------------------------------------------------------------------
static int ensure_pipe(dbv_ctx* c) {
  if (c->pipe_ready) return DBV_OK;
  DBV_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  DBV_CUDA(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
  DBV_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  int r;
  for (int i = 0; i < 2; ++i) {
    if ((r = dev_alloc(c, &c->stage_in[i], (size_t)c->chunk * STAMP_ELTS * 8, false))) return r;
    if ((r = dev_alloc(c, (void**)&c->stage_x[i], (size_t)c->chunk * STAMP_ELTS * 4, false))) return r;
    if ((r = dev_alloc(c, (void**)&c->stage_mean[i], (size_t)c->chunk * STAMP_ELTS * 4, false))) return r;
    if ((r = dev_alloc(c, (void**)&c->stage_std[i], (size_t)c->chunk * STAMP_ELTS * 4, false))) return r;
    if ((r = dev_alloc(c, (void**)&c->stage_z[i], (size_t)c->chunk * LAT * 4, false))) return r;
    if ((r = dev_alloc(c, (void**)&c->stage_eps[i], (size_t)c->chunk * LAT * 4, false))) return r;
    DBV_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    DBV_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
    DBV_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
  }
  c->pipe_ready = true;
  return DBV_OK;
}

// Piece schedule of the H2D / compute / D2H pipeline of dbv_deblend_host: (first stamp, count) of every piece.
// Measured on B200 (tools/e2e_breakdown.py, pcie_probe.py): a piece of n stamps costs ~0.35 ms + 1.95 us * n of compute,
// 1.5 us * n of H2D and 1.46 us * n of D2H; the first H2D and the last D2H cannot overlap anything.  So: start small and let
// the pieces GROW as fast as the copies keep up with the compute (H2D of piece k+1 <= compute of piece k: n' <= 1.25 n + 224),
// cap them, and finish with one short piece.  The old fixed-size schedule (DBV_HOST_PIECE=n) stalled ~0.7 ms on its second
// piece and paid three short tail pieces.
static std::vector<std::pair<int64_t, long long>> host_schedule(int64_t B, long long chunk) {
  std::vector<std::pair<int64_t, long long>> sched;
  if (const char* e = dbv_env("DBV_HOST_PIECE")) {
    int64_t b0 = 0;
    const long long piece = std::max<long long>(1, std::min<long long>(chunk, atoll(e)));
    const long long q = std::max<long long>(piece / 4, 1);
    if (B > 2 * piece) { sched.push_back({0, q}); b0 = q; }
    while (B - b0 > piece + q) { sched.push_back({b0, piece}); b0 += piece; }
    while (b0 < B) {
      long long rem = B - b0, nb = rem;
      if (B > 2 * piece && rem > q) nb = std::max<long long>((rem + 1) / 2, q);
      nb = std::min<long long>(nb, piece);
      sched.push_back({b0, nb});
      b0 += nb;
    }
  } else {
    const long long cap = std::min<long long>(chunk, 1376), first = std::min<long long>(chunk, 256), last = first;
    int64_t b0 = 0;
    long long n = first;
    while (B - b0 > 0) {
      const long long rem = B - b0;
      long long nb;
      if (rem <= first + first / 4) nb = rem;   // the last, short piece
      else if (rem - n >= last) nb = n;         // ramp / steady state: a tail piece still fits behind it
      else nb = rem - last;                     // the piece before the last one
      nb = std::min<long long>(nb, chunk);
      sched.push_back({b0, nb});
      b0 += nb;
      n = std::min<long long>(cap, ((5 * n / 4 + 224) / 32) * 32);
    }
  }
  return sched;
}

extern "C" int64_t dbv_host_schedule(int64_t B, int64_t chunk, int64_t* counts, int64_t max_pieces) {
  if (B < 0 || chunk <= 0 || (max_pieces > 0 && !counts)) return DBV_ERR_INVALID;
  const auto sched = host_schedule(B, chunk);
  for (size_t i = 0; i < sched.size() && (int64_t)i < max_pieces; ++i) counts[i] = sched[i].second;
  return (int64_t)sched.size();
}

extern "C" int dbv_deblend_host(dbv_ctx* c, const void* x_host, int x_dtype, int64_t B, const float* eps_host, uint64_t seed,
                                int sample, float* mean_host, float* stddev_host, float* z_host, float* mean_dev,
                                float* stddev_dev) {
  int r = check_ready(c, "dbv_deblend_host");
  if (r) return r;
  DBV_REQUIRE(B == 0 || (x_host && mean_host), "dbv_deblend_host: null buffer");
  DBV_REQUIRE(x_dtype == DBV_F32 || x_dtype == DBV_F64, "dbv_deblend_host: bad dtype %d", x_dtype);
  DBV_REQUIRE(B >= 0, "dbv_deblend_host: negative B");
  if ((r = ensure_pipe(c))) return r;
  const long long before = g_launches.load();
  const size_t esz = x_dtype == DBV_F64 ? 8 : 4;
  const std::vector<std::pair<int64_t, long long>> sched = host_schedule(B, c->chunk);
  // Pageable input (an ordinary numpy array): host threads stage it — converting float64 to float32 on the way, which also
  // halves the PCIe bytes — into pinned memory.  Pinned (registered) input goes to the device as it is.
  bool stage = false;
  if (B > 0) {
    cudaPointerAttributes at{};
    const cudaError_t pe = cudaPointerGetAttributes(&at, x_host);
    if (pe != cudaSuccess) cudaGetLastError();  // unregistered memory is reported as an error by old runtimes
    if (pe == cudaSuccess && at.type == cudaMemoryTypeDevice)
      return fail(DBV_ERR_INVALID, "dbv_deblend_host: x_host is a device pointer (device-resident stamps go through dbv_deblend)");
    stage = !(pe == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged));
  }
  if (stage) {
    long long mx = 0;
    for (const auto& pc : sched) mx = std::max(mx, pc.second);
    const size_t need = (size_t)mx * STAMP_ELTS;
    if (need > c->hstage_elems) {
      for (int i = 0; i < 2; ++i) {
        if (c->hstage[i]) cudaFreeHost(c->hstage[i]);
        c->hstage[i] = nullptr;
      }
      c->hstage_elems = 0;
      for (int i = 0; i < 2; ++i) DBV_CUDA(cudaHostAlloc((void**)&c->hstage[i], need * sizeof(float), cudaHostAllocDefault));
      c->hstage_elems = need;
    }
    if (!c->hpool) c->hpool = new HostStagePool(host_stage_default_threads());
  }
  // the activation buffers may still be in use by a dbv_deblend / dbv_encode / dbv_decode call enqueued on a caller's stream
  if (c->ev_user) DBV_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_user, 0));
  // on an error in the middle of the pipeline the copies already enqueued may still be writing the caller's host buffers:
  // drain the three streams before returning
  auto drain = [&](int rc) {
    cudaStreamSynchronize(c->s_d2h);
    cudaStreamSynchronize(c->s_comp);
    cudaStreamSynchronize(c->s_h2d);
    return rc;
  };
#define DBV_PIPE(expr)                                                                                            \
  do {                                                                                                            \
    cudaError_t e_ = (expr);                                                                                      \
    if (e_ != cudaSuccess) return drain(fail(DBV_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)));             \
  } while (0)
  for (int k = 0; k < (int)sched.size(); ++k) {
    const int s = k & 1;
    const int64_t b0 = sched[k].first;
    const long long nb = sched[k].second;
    const size_t n = (size_t)nb * STAMP_ELTS;
    // slot s is free for new input once the compute that used it (chunk k-2) is done, and its
    // outputs are free once the D2H of chunk k-2 is done
    if (k >= 2) {
      DBV_PIPE(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[s], 0));
      DBV_PIPE(cudaStreamWaitEvent(c->s_comp, c->ev_out[s], 0));
    }
    if (stage) {
      // the H2D copy that last read this staging slot (piece k-2) must have finished before the slot is refilled
      if (k >= 2) DBV_PIPE(cudaEventSynchronize(c->ev_in[s]));
      c->hpool->convert((const char*)x_host + (size_t)b0 * STAMP_ELTS * esz, x_dtype == DBV_F64, c->hstage[s], n);
      DBV_PIPE(cudaMemcpyAsync(c->stage_x[s], c->hstage[s], n * 4, cudaMemcpyHostToDevice, c->s_h2d));
    } else {
      void* dst = x_dtype == DBV_F64 ? c->stage_in[s] : (void*)c->stage_x[s];
      DBV_PIPE(cudaMemcpyAsync(dst, (const char*)x_host + (size_t)b0 * STAMP_ELTS * esz, n * esz, cudaMemcpyHostToDevice, c->s_h2d));
    }
    if (eps_host)
      DBV_PIPE(cudaMemcpyAsync(c->stage_eps[s], eps_host + b0 * LAT, (size_t)nb * LAT * 4, cudaMemcpyHostToDevice, c->s_h2d));
    DBV_PIPE(cudaEventRecord(c->ev_in[s], c->s_h2d));
    DBV_PIPE(cudaStreamWaitEvent(c->s_comp, c->ev_in[s], 0));
    if (!stage && x_dtype == DBV_F64 && (r = launch_cast_f64_f32((const double*)c->stage_in[s], c->stage_x[s], (long long)n, c->s_comp))) return drain(r);
    // outputs: the caller's resident device buffers when given, else the double-buffered staging slots
    float* d_mean = mean_dev ? mean_dev + (size_t)b0 * STAMP_ELTS : c->stage_mean[s];
    float* d_std = stddev_dev ? stddev_dev + (size_t)b0 * STAMP_ELTS : (stddev_host ? c->stage_std[s] : nullptr);
    r = run_chunk(c, c->stage_x[s], nb, eps_host ? c->stage_eps[s] : nullptr, seed, sample, b0, nullptr, c->stage_z[s], nullptr,
                  nullptr, d_mean, d_std, true, true, true, c->s_comp);
    if (r) return drain(r);
    DBV_PIPE(cudaEventRecord(c->ev_comp[s], c->s_comp));
    DBV_PIPE(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[s], 0));
    DBV_PIPE(cudaMemcpyAsync(mean_host + (size_t)b0 * STAMP_ELTS, d_mean, n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    if (stddev_host)
      DBV_PIPE(cudaMemcpyAsync(stddev_host + (size_t)b0 * STAMP_ELTS, d_std, n * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    if (z_host) DBV_PIPE(cudaMemcpyAsync(z_host + b0 * LAT, c->stage_z[s], (size_t)nb * LAT * 4, cudaMemcpyDeviceToHost, c->s_d2h));
    DBV_PIPE(cudaEventRecord(c->ev_out[s], c->s_d2h));
  }
  DBV_CUDA(cudaStreamSynchronize(c->s_d2h));
  DBV_CUDA(cudaStreamSynchronize(c->s_comp));
  DBV_CUDA(cudaStreamSynchronize(c->s_h2d));
#undef DBV_PIPE
  c->launches += g_launches.load() - before;
  return DBV_OK;
}

extern "C" int dbv_fp16_overflow(dbv_ctx* c, int reset) {
  DBV_REQUIRE(c, "dbv_fp16_overflow: null ctx");
  if (!c->ovf_host) return 0;
  const int v = *reinterpret_cast<volatile int*>(c->ovf_host);
  if (reset) *reinterpret_cast<volatile int*>(c->ovf_host) = 0;
  return v != 0 ? 1 : 0;
}

// ---- introspection -----------------------------------------------------------------------------------
extern "C" int dbv_set_profiling(dbv_ctx* c, int enabled) {
  DBV_REQUIRE(c, "dbv_set_profiling: null ctx");
  c->profiling = enabled != 0;
  c->prof_n = 0;
  c->prof_calls = 0;
  return DBV_OK;
}

extern "C" int dbv_layer_times(dbv_ctx* c, int max_layers, float* ms_out, char* names_out) {
  DBV_REQUIRE(c && ms_out && names_out, "dbv_layer_times: null argument");
  // average per call since profiling was switched on: sum over all recorded chunks / number of calls; layers in order of
  // first appearance; "start" marks open a new chunk
  std::vector<std::string> names;
  std::vector<float> tot;
  for (int i = 1; i < c->prof_n; ++i) {
    if (c->prof_names[i] == "start") continue;
    float ms = 0.f;
    cudaError_t e = cudaEventElapsedTime(&ms, c->prof_ev[i - 1], c->prof_ev[i]);
    if (e != cudaSuccess) return fail(DBV_ERR_CUDA, "dbv_layer_times: %s", cudaGetErrorString(e));
    size_t k = 0;
    for (; k < names.size(); ++k)
      if (names[k] == c->prof_names[i]) break;
    if (k == names.size()) { names.push_back(c->prof_names[i]); tot.push_back(0.f); }
    tot[k] += ms;
  }
  int n = 0;
  const float calls = (float)(c->prof_calls > 0 ? c->prof_calls : 1);
  for (; n < (int)names.size() && n < max_layers; ++n) {
    ms_out[n] = tot[n] / calls;
    strncpy(names_out + 32 * n, names[n].c_str(), 31);
    names_out[32 * n + 31] = 0;
  }
  return n;
}

extern "C" int dbv_layer_kernel(dbv_ctx* c, const char* name, char* out, int out_bytes) {
  // which __global__ function (as the ncu launch list names it) runs this layer under the ctx's precision and tuned plan
  int r = check_ready(c, "dbv_layer_kernel");
  if (r) return r;
  DBV_REQUIRE(name && out && out_bytes >= 48, "dbv_layer_kernel: bad argument");
  if (!strcmp(name, "enc_bn_pack")) { snprintf(out, out_bytes, "bn_pack8_kernel"); return DBV_OK; }
  if (!strcmp(name, "latent")) { snprintf(out, out_bytes, "latent_kernel"); return DBV_OK; }
  for (int li = 0; li < kNumLayers; ++li) {
    if (strcmp(name, kLayers[li].name)) continue;
    const LayerRt& R = c->rt[li];
    if (R.has_halo) snprintf(out, out_bytes, "%s<%d, %d>", R.halo.pair ? "tc_halo2_kernel" : "tc_halo_kernel", kTc[li].CBK, kTc[li].NT);
    else if (R.has_pairh) snprintf(out, out_bytes, "tc_pairh_kernel<%d>", kTc[li].NT);
    else if (R.has_pair) snprintf(out, out_bytes, "tc_pair_kernel<%d, %d>", kTc[li].CBK, kTc[li].NT);
    else if (R.has_tc) { const TcGeom G = tc_geom(c->precision, li); snprintf(out, out_bytes, "tc_conv_kernel<%d, %d>", G.CBK, G.NT); }
    else snprintf(out, out_bytes, "simt_kernel");
    return DBV_OK;
  }
  return fail(DBV_ERR_INVALID, "dbv_layer_kernel: unknown layer '%s'", name);
}

extern "C" int dbv_debug_activation(dbv_ctx* c, const char* name, int64_t B, float* out, void* stream) {
  int r = check_ready(c, "dbv_debug_activation");
  if (r) return r;
  DBV_REQUIRE(name && out && B >= 0 && B <= c->chunk, "dbv_debug_activation: bad argument");
  for (int li = 0; li < kNumLayers; ++li)
    if (!strcmp(name, kLayers[li].name)) {
      if (li == I_HEAD) return fail(DBV_ERR_INVALID, "dbv_debug_activation: the head writes the caller's buffers");
      return launch_act_to_f32(c->rt[li].ospec, B, out, (cudaStream_t)stream);
    }
  return fail(DBV_ERR_INVALID, "dbv_debug_activation: unknown layer '%s'", name);
}
