// libpf_b200.so -- C ABI (include/pf_b200.h) and forward orchestration of the B200-native PerspectiveFields
// inference engine.  One engine per device; pf_forward enqueues the whole graph of
// perspective2d/perspectivefields.py:223-272 on the caller's stream.
#include "../../include/pf_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <functional>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "tma_host.cuh"
#include "attention_mma.cuh"
#include "layers.cuh"
#include "prepost.cuh"
#include "comm.cuh"
#include "jpeg.cuh"

using namespace pf;

// ----------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(expr)                                                                                    \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) return fail(PF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
// PF_SYNC_DEBUG=1 in the environment: synchronise after every launch so that a device fault is reported at the launch that
// caused it (debugging aid; never set in production)
static char g_crumb[96] = "";   // name of the last debug tap taken (breadcrumb for the error text)
static bool sync_debug() {
  static int v = -1;
  if (v < 0) v = getenv("PF_SYNC_DEBUG") ? 1 : 0;
  return v == 1;
}
// pf_profile_kernels_*: a CUDA-event pair around EVERY launch of the forward graph (in-pipeline time per kernel, bench.py's
// "per_kernel" table).  Off by default: the event records cost a few percent, so bench.py uses a separate pass for it.
// The state lives in the engine; the launch macro reaches it through a thread-local pointer that pf_forward sets for the
// duration of the call (operator entry points run with it unset).
struct KernelProf {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { const char* expr; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t next() { return used < pool.size() ? pool[used++] : nullptr; }
};
static thread_local KernelProf* tl_kp = nullptr;
#define LAUNCHED(expr)                                                                              \
  do {                                                                                              \
    KernelProf* kp__ = tl_kp;                                                                       \
    cudaEvent_t ka__ = (kp__ && kp__->on) ? kp__->next() : nullptr;                                 \
    if (ka__) cudaEventRecord(ka__, kp__->st);                                                      \
    cudaError_t e__ = (expr);                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
    if (ka__) {                                                                                     \
      cudaEvent_t kb__ = kp__->next();                                                              \
      if (kb__) { cudaEventRecord(kb__, kp__->st); kp__->recs.push_back({#expr, ka__, kb__}); }     \
    }                                                                                               \
    if (e__ == cudaSuccess && sync_debug()) e__ = cudaDeviceSynchronize();                          \
    if (e__ != cudaSuccess) return fail(PF_ERR_CUDA, "%s: %s (%s:%d, after tap '%s')", #expr, cudaGetErrorString(e__), __FILE__, __LINE__, g_crumb); \
  } while (0)
// NVTX range per section of the forward graph (visible in Nsight Systems / ncu --nvtx; a few ns when no tool is attached)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define TRY(expr)                \
  do {                           \
    int r__ = (expr);            \
    if (r__ != PF_OK) return r__; \
  } while (0)

// ----------------------------------------------------------------------------------------------- model constants
static const int kMitDims[4] = {64, 128, 320, 512};
static const int kMitHeads[4] = {1, 2, 5, 8};
static const int kMitDepths[4] = {3, 4, 18, 3};
static const int kMitSr[4] = {8, 4, 2, 1};
static const int kMitRes[4] = {80, 40, 20, 10};
static const int kCnxDims[4] = {96, 192, 384, 768};
static const int kCnxDepths[4] = {3, 3, 9, 3};

struct WeightRef { const void* p; long long numel; int dtype; };
struct GemmW { const __nv_bfloat16* hi = nullptr; const __nv_bfloat16* lo = nullptr; const float* b = nullptr; };
struct LnW { const float* w = nullptr; const float* b = nullptr; };

struct MitBlockW { LnW ln1, srln, ln2; GemmW q, sr, kv, proj, fc1, fc2; const float* dw_w; const float* dw_b; };
struct CnxBlockW { const float* dw_w; const float* dw_b; LnW ln; GemmW pw1, pw2; const float* gamma; };

struct Arena {
  char* base = nullptr;
  long long cap = 0, off = 0, peak = 0;
  bool dry = false, keep = false;  // keep: debug mode, never recycle
  void* alloc(long long bytes) {
    off = (off + 255) & ~255LL;
    void* p = dry ? nullptr : base + off;
    off += bytes;
    if (off > peak) peak = off;
    return p;
  }
  float* f(long long n) { return (float*)alloc(n * 4); }
  long long mark() const { return off; }
  void release(long long m) { if (!keep) off = m; }
};

struct pf_engine {
  int device = 0;
  pf_model_desc desc{};
  bool finalized = false;
  std::unordered_map<std::string, WeightRef> weights;
  // resolved weights
  const float *embed1_w, *embed1_b, *llenc_w, *llenc_b;
  LnW embed_ln[4], stage_norm[4];
  GemmW embed[4];  // [1..3] used
  GemmW embed1g, llencg;  // 7x7 stems as [64][160] GEMMs (tensor-core path)
  std::vector<MitBlockW> blocks[4];
  GemmW proc[4];   // composed linear_c{l} o linear_c{l}_proc, both heads side by side (N = 512), index lvl-1
  GemmW rcu[4][2][2];  // [fusion-1][unit-1][conv-1], grouped over the two heads
  GemmW conv0, conv1;
  GemmW conv1p;                       // conv_fuse_conv1 composed with the x2 upsample in front of it: 4 phases x 32 outputs per head
  const float *conv1f_w, *conv1f_b;   // plain fp32 conv_fuse_conv1 [head][tap][ci][o] / bias, for the border-ring kernel
  bool use_attn_tc = true;            // option "attn_tc": attention core on tcgen05 / TMEM (attention_tc.cuh); 0 = warp-level mma.sync kernel
  bool use_fork = false;              // option "fork": the spatial-reduction branch of a MiT block (sr conv -> LayerNorm -> kv) runs on a second
                                      // stream next to the q projection (both only depend on LayerNorm 1; their grids leave SMs idle: 50-75 tiles).
                                      // Default OFF: measured 1 % slower (the persistent kernels of the two streams compete for SMs and the
                                      // event waits break the programmatic-dependent-launch chain; profiles/r02_notes.md)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_pair = true;               // option "pair": GEMM-mode launches with enough tiles run on CTA pairs (gemm2_tma.cuh, cta_group::2)
  bool use_dwln = false;              // option "dw_ln": ConvNeXt depthwise 7x7 fused with the LayerNorm that follows it
  bool use_pdl = true;                // option "pdl": programmatic dependent launch of the graph's kernels (common.cuh)
  bool decode_only = false;           // option "decode_only": classification heads return decoded fields, logits are never written
  bool use_attn_split = true;         // option "attn_split": q / kv leave their GEMMs as split planes (0 = fp32, split inside the attention kernel)
  bool use_phase = true;              // option "phase_conv1": 0 = materialise the upsampled tensor and run conv1 at 320x320
  const float *pred_g_w, *pred_g_b, *pred_l_w, *pred_l_b;
  const float *pn_stem_w, *pn_stem_b;
  LnW pn_stem_ln, pn_ds_ln[4], pn_norm;
  GemmW pn_ds[4];
  std::vector<CnxBlockW> pn_blocks[4];
  const float *pn_head_w, *pn_head_b;
  // Pillow resample tables, cached per input size in one device slab owned by the engine (bump allocation; built on the host
  // into a pinned mirror of the slab and copied with cudaMemcpyAsync on the caller's stream: no allocation and no
  // synchronising copy inside pf_forward)
  struct DevTable { int ksize; int* bounds; int* coeffs; };
  std::map<int, DevTable> tables;
  char* table_dev = nullptr;
  char* table_host = nullptr;       // pinned
  long long table_off = 0;
  KernelProf kp;                    // pf_profile_kernels_*
  // per-launch profiling of the GEMM engine (bench.py roofline leg): CUDA events on the launch stream
  // tensor maps are pure functions of (pointer, shape, box): cached across calls (the arena hands out the same addresses for the
  // same batch size), which takes cuTensorMapEncodeTiled (~5 us each, ~1800 per forward) off the launch path
  struct MapKey {
    const void* base; long long d0, d1, d2; int kind, box, kb;
    bool operator==(const MapKey& o) const { return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && kind == o.kind && box == o.box && kb == o.kb; }
  };
  struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
      size_t h = std::hash<const void*>()(k.base);
      for (long long v : {k.d0, k.d1, k.d2, (long long)k.kind, (long long)k.box, (long long)k.kb}) h = h * 1000003u ^ std::hash<long long>()(v);
      return h;
    }
  };
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> map_cache;
  bool use_stem_tc = true;    // 7x7 stems as patch gather + TMA GEMM (option "stem_tc"; 0 = fp32 CUDA-core direct convolution)
  bool use_attn_mma = true;   // tensor-core attention core (option "attn_mma"; 0 = CUDA-core fp32 kernel)
  int sm_count = 148;
  bool profile = false;
  struct ProfRec { cudaEvent_t a, b; double flops; int cfg; int M, N, K, KH, stride, groups, Cin; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;   // events are created once and recycled: no create/destroy inside a timed region
  // debug taps
  bool debug = false;
  std::vector<std::pair<std::string, std::pair<const float*, long long>>> taps;
};

// ----------------------------------------------------------------------------------------------- weight lookup
static int get_w(pf_engine* e, const std::string& name, int dtype, long long numel, const void** out) {
  auto it = e->weights.find(name);
  if (it == e->weights.end()) return fail(PF_ERR_WEIGHT, "missing weight '%s'", name.c_str());
  if (it->second.dtype != dtype) return fail(PF_ERR_WEIGHT, "weight '%s': wrong dtype", name.c_str());
  if (it->second.numel != numel) return fail(PF_ERR_WEIGHT, "weight '%s': numel %lld, expected %lld", name.c_str(), it->second.numel, numel);
  *out = it->second.p;
  return PF_OK;
}
static int get_f(pf_engine* e, const std::string& n, long long numel, const float** out) { return get_w(e, n, PF_F32, numel, (const void**)out); }
static int get_gemm(pf_engine* e, const std::string& n, long long N, long long K, long long nbias, GemmW* g, int groups = 1) {
  TRY(get_w(e, n + ".whi", PF_BF16, groups * N * K, (const void**)&g->hi));
  TRY(get_w(e, n + ".wlo", PF_BF16, groups * N * K, (const void**)&g->lo));
  TRY(get_f(e, n + ".b", groups * nbias, &g->b));
  return PF_OK;
}
static int get_ln(pf_engine* e, const std::string& n, int C, LnW* l) {
  TRY(get_f(e, n + ".w", C, &l->w));
  TRY(get_f(e, n + ".b", C, &l->b));
  return PF_OK;
}

static int resolve_weights(pf_engine* e) {
  char nm[128];
  TRY(get_f(e, "embed1.w", 147 * 64, &e->embed1_w));
  TRY(get_f(e, "embed1.b", 64, &e->embed1_b));
  TRY(get_f(e, "llenc.w", 147 * 64, &e->llenc_w));
  TRY(get_f(e, "llenc.b", 64, &e->llenc_b));
  TRY(get_gemm(e, "embed1g", 64, 160, 64, &e->embed1g));
  TRY(get_gemm(e, "llencg", 64, 160, 64, &e->llencg));
  for (int s = 0; s < 4; ++s) {
    const int C = kMitDims[s];
    snprintf(nm, sizeof nm, "embed%d.ln", s + 1);
    TRY(get_ln(e, nm, C, &e->embed_ln[s]));
    if (s > 0) {
      snprintf(nm, sizeof nm, "embed%d", s + 1);
      TRY(get_gemm(e, nm, C, 9LL * kMitDims[s - 1], C, &e->embed[s]));
    }
    e->blocks[s].resize(kMitDepths[s]);
    for (int i = 0; i < kMitDepths[s]; ++i) {
      MitBlockW& b = e->blocks[s][i];
      char p[64];
      snprintf(p, sizeof p, "s%d.b%d.", s + 1, i);
      std::string P(p);
      TRY(get_ln(e, P + "ln1", C, &b.ln1));
      TRY(get_gemm(e, P + "q", C, C, C, &b.q));
      if (kMitSr[s] > 1) {
        TRY(get_gemm(e, P + "sr", C, (long long)kMitSr[s] * kMitSr[s] * C, C, &b.sr));
        TRY(get_ln(e, P + "srln", C, &b.srln));
      }
      TRY(get_gemm(e, P + "kv", 2 * C, C, 2 * C, &b.kv));
      TRY(get_gemm(e, P + "proj", C, C, C, &b.proj));
      TRY(get_ln(e, P + "ln2", C, &b.ln2));
      TRY(get_gemm(e, P + "fc1", 4 * C, C, 4 * C, &b.fc1));
      TRY(get_f(e, P + "dw.w", 9LL * 4 * C, &b.dw_w));
      TRY(get_f(e, P + "dw.b", 4 * C, &b.dw_b));
      TRY(get_gemm(e, P + "fc2", C, 4 * C, C, &b.fc2));
    }
    snprintf(nm, sizeof nm, "s%d.norm", s + 1);
    TRY(get_ln(e, nm, C, &e->stage_norm[s]));
  }
  for (int l = 0; l < 4; ++l) {
    snprintf(nm, sizeof nm, "head.proc%d", l + 1);
    TRY(get_gemm(e, nm, 512, 9LL * kMitDims[l], 9 * 512, &e->proc[l]));
  }
  for (int f = 0; f < 4; ++f)
    for (int u = 0; u < 2; ++u) {
      if (f == 3 && u == 0) continue;  // fusion4 has resConfUnit2 only (gravity_head.py:102)
      for (int c = 0; c < 2; ++c) {
        snprintf(nm, sizeof nm, "head.f%d.u%d.c%d", f + 1, u + 1, c + 1);
        TRY(get_gemm(e, nm, 256, 2304, 256, &e->rcu[f][u][c], 2));
      }
    }
  TRY(get_gemm(e, "head.conv0", 64, 9 * 320, 64, &e->conv0, 2));
  TRY(get_gemm(e, "head.conv1", 32, 9 * 64, 32, &e->conv1, 2));
  TRY(get_gemm(e, "head.conv1p", 128, 9 * 64, 128, &e->conv1p, 2));
  TRY(get_f(e, "head.conv1f.w", 2LL * 9 * 64 * 32, &e->conv1f_w));
  TRY(get_f(e, "head.conv1f.b", 64, &e->conv1f_b));
  TRY(get_f(e, "head.pred_g.w", 32LL * e->desc.gravity_classes, &e->pred_g_w));
  TRY(get_f(e, "head.pred_g.b", e->desc.gravity_classes, &e->pred_g_b));
  TRY(get_f(e, "head.pred_l.w", 32LL * e->desc.latitude_classes, &e->pred_l_w));
  TRY(get_f(e, "head.pred_l.b", e->desc.latitude_classes, &e->pred_l_b));
  if (e->desc.param_net != PF_PARAM_NONE) {
    TRY(get_f(e, "pn.stem.w", 48 * 96, &e->pn_stem_w));
    TRY(get_f(e, "pn.stem.b", 96, &e->pn_stem_b));
    TRY(get_ln(e, "pn.stem.ln", 96, &e->pn_stem_ln));
    for (int k = 1; k < 4; ++k) {
      snprintf(nm, sizeof nm, "pn.ds%d.ln", k);
      TRY(get_ln(e, nm, kCnxDims[k - 1], &e->pn_ds_ln[k]));
      snprintf(nm, sizeof nm, "pn.ds%d", k);
      TRY(get_gemm(e, nm, kCnxDims[k], 4LL * kCnxDims[k - 1], kCnxDims[k], &e->pn_ds[k]));
    }
    for (int s = 0; s < 4; ++s) {
      const int C = kCnxDims[s];
      e->pn_blocks[s].resize(kCnxDepths[s]);
      for (int j = 0; j < kCnxDepths[s]; ++j) {
        CnxBlockW& b = e->pn_blocks[s][j];
        char p[64];
        snprintf(p, sizeof p, "pn.s%d.b%d.", s, j);
        std::string P(p);
        TRY(get_f(e, P + "dw.w", 49LL * C, &b.dw_w));
        TRY(get_f(e, P + "dw.b", C, &b.dw_b));
        TRY(get_ln(e, P + "ln", C, &b.ln));
        TRY(get_gemm(e, P + "pw1", 4 * C, C, 4 * C, &b.pw1));
        TRY(get_gemm(e, P + "pw2", C, 4 * C, C, &b.pw2));
        TRY(get_f(e, P + "gamma", C, &b.gamma));
      }
    }
    TRY(get_ln(e, "pn.norm", 768, &e->pn_norm));
    TRY(get_f(e, "pn.head.w", 5 * 768, &e->pn_head_w));
    TRY(get_f(e, "pn.head.b", 5, &e->pn_head_b));
  }
  return PF_OK;
}

// ----------------------------------------------------------------------------------------------- op helpers
struct Fwd {
  pf_engine* e;
  Arena ar;
  cudaStream_t st;
  bool dry;
  int n;

  // debug taps: snapshot the tensor into a private buffer (many intermediates are updated in place later)
  int tap(const char* name, const float* p, long long numel) {
    if (!e->debug) return PF_OK;
    float* cp = ar.f(numel);
    if (dry) return PF_OK;
    snprintf(g_crumb, sizeof g_crumb, "%s", name);
    if (sync_debug()) fprintf(stderr, "[pf tap] %s cp=%p (+%lld of cap %lld) p=%p numel=%lld\n", name, (void*)cp, (long long)((char*)cp - ar.base), ar.cap, (const void*)p, numel);
    CU(cudaMemcpyAsync(cp, p, numel * 4, cudaMemcpyDeviceToDevice, st));
    if (sync_debug()) CU(cudaDeviceSynchronize());
    e->taps.push_back({name, {cp, numel}});
    return PF_OK;
  }
  int tapf(const float* p, long long numel, const char* fmt, ...) {
    if (!e->debug) return PF_OK;
    char buf[96];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return tap(buf, p, numel);
  }

  // ------------------------------------------------------------------------------------------ TMA engine helpers
  SplitT salloc(long long pixels, int ld) {
    SplitT t;
    t.hi = (__nv_bfloat16*)ar.alloc(pixels * ld * 2);
    t.lo = (__nv_bfloat16*)ar.alloc(pixels * ld * 2);
    t.ld = ld;
    return t;
  }
  int tap_split(const char* name, const SplitT& t, long long numel) {
    if (!e->debug) return PF_OK;
    float* cp = ar.f(numel);
    if (dry) return PF_OK;
    snprintf(g_crumb, sizeof g_crumb, "%s", name);
    LAUNCHED((merge_split_kernel<<<(unsigned)cdivl(numel, 256), 256, 0, st>>>(t.hi, t.lo, cp, numel), cudaGetLastError()));
    e->taps.push_back({name, {cp, numel}});
    return PF_OK;
  }
  // (two names so that the per-kernel profile separates the GEMM-mode and halo-mode launches)
  static cudaError_t gemm_tma_gemm_mode(const TmaMaps& maps, const TmaGemmParams& p, int bn, int kb, int sms, cudaStream_t st, const PredTail* pred) {
    return gemm_tma_launch(MODE_GEMM, maps, p, bn, kb, sms, st, pred);
  }
  static cudaError_t gemm_tma_halo_mode(const TmaMaps& maps, const TmaGemmParams& p, int bn, int kb, int sms, cudaStream_t st, const PredTail* pred) {
    return gemm_tma_launch(MODE_HALO, maps, p, bn, kb, sms, st, pred);
  }
  int launch_tma(int mode, const TmaMaps& maps, const TmaGemmParams& p, const PredTail* pred = nullptr) {
    const int bn = mode == MODE_GEMM ? tma_pick_bn_gemm(p.M, p.N, p.K, e->sm_count) : tma_pick_bn(p.N, mode), kb = tma_pick_kb(bn, p.K, mode);
    if (e->profile) {
      pf_engine::ProfRec r{};
      for (cudaEvent_t* ev : {&r.a, &r.b}) {
        if (e->ev_pool.empty()) { CU(cudaEventCreate(ev)); }
        else { *ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
      }
      const double Mrows = mode == MODE_GEMM ? (double)p.M : (double)p.B * p.H * p.W;
      r.flops = 2.0 * Mrows * (double)p.N * (double)p.K * (double)p.groups;
      r.cfg = mode == MODE_GEMM ? 5 : 6;
      r.M = (int)Mrows; r.N = p.N; r.K = p.K; r.KH = mode == MODE_GEMM ? 1 : 3; r.stride = 1; r.groups = p.groups; r.Cin = p.Cin;
      CU(cudaEventRecord(r.a, st));
      if (mode == MODE_GEMM) LAUNCHED(gemm_tma_gemm_mode(maps, p, bn, kb, e->sm_count, st, pred));
      else LAUNCHED(gemm_tma_halo_mode(maps, p, bn, kb, e->sm_count, st, pred));
      CU(cudaEventRecord(r.b, st));
      e->prof.push_back(r);
      return PF_OK;
    }
    if (mode == MODE_GEMM) LAUNCHED(gemm_tma_gemm_mode(maps, p, bn, kb, e->sm_count, st, pred));
    else LAUNCHED(gemm_tma_halo_mode(maps, p, bn, kb, e->sm_count, st, pred));
    return PF_OK;
  }
  struct Epi {   // epilogue options of one TMA GEMM / conv
    float* C = nullptr; int ldc = 0, c_coff = 0, c_gcoff = 0;
    SplitT S; int s_coff = 0, s_gcoff = 0, split_relu = 0;
    int act = 0; const float* gamma = nullptr;
    const float* res = nullptr; int ldr = 0, r_coff = 0, r_gcoff = 0, res_relu = 0;
    const float* res2 = nullptr; int ldr2 = 0, r2_coff = 0, r2_gcoff = 0;
    int bias_mode = 1;
    int phase4 = 0;   // halo mode, N = 128: columns are 4 output phases x 32 channels of a 2H x 2W output (TmaGemmParams::phase4)
  };
  static void fill_epi(TmaGemmParams& p, const GemmW& w, const Epi& o, int bias_gstride) {
    p.bias = w.b; p.bias_mode = w.b ? o.bias_mode : 0; p.bias_gstride = bias_gstride;
    p.act = o.act; p.gamma = o.gamma;
    p.res = o.res; p.ldr = o.ldr; p.r_coff = o.r_coff; p.r_gcoff = o.r_gcoff; p.res_relu = o.res_relu;
    p.res2 = o.res2; p.ldr2 = o.ldr2; p.r2_coff = o.r2_coff; p.r2_gcoff = o.r2_gcoff;
    p.C = o.C; p.ldc = o.ldc; p.c_coff = o.c_coff; p.c_gcoff = o.c_gcoff;
    p.Shi = o.S.hi; p.Slo = o.S.lo; p.lds = o.S.ld; p.s_coff = o.s_coff; p.s_gcoff = o.s_gcoff; p.split_relu = o.split_relu;
    p.phase4 = o.phase4;
  }
  // cached tensor-map constructors
  template <class F>
  const char* cached_map(CUtensorMap* out, const pf_engine::MapKey& key, F&& make) {
    auto it = e->map_cache.find(key);
    if (it != e->map_cache.end()) { *out = it->second; return nullptr; }
    const char* msg = make(out);
    if (!msg) {
      if (e->map_cache.size() > 20000) e->map_cache.clear();
      e->map_cache.emplace(key, *out);
    }
    return msg;
  }
  const char* map2d(CUtensorMap* m, const void* base, long long cols, long long rows, long long ld, int box_rows, int kb) {
    return cached_map(m, pf_engine::MapKey{base, cols, rows, ld, 0, box_rows, kb}, [&](CUtensorMap* o) { return tma_map_2d(o, base, cols, rows, ld, box_rows, kb); });
  }
  const char* map_tile32(CUtensorMap* m, const void* base, long long rows, long long ld, bool f32) {
    return cached_map(m, pf_engine::MapKey{base, rows, ld, 0, f32 ? 1 : 2, 32, 0}, [&](CUtensorMap* o) { return tma_map_tile32(o, base, rows, ld, f32); });
  }
  const char* map_halo(CUtensorMap* m, const void* base, int B, int H, int W, int ld) {
    return cached_map(m, pf_engine::MapKey{base, ((long long)B << 32) | (unsigned)H, W, ld, 3, 0, 0}, [&](CUtensorMap* o) { return tma_map_halo(o, base, B, H, W, ld); });
  }
  // C[M, N] = A[M, K] W^T : A = split planes with row pitch A.ld, first channel a_c0
  int tgemm(const SplitT& A, long long M, int K, int a_c0, const GemmW& w, int N, const Epi& o) {
    if (dry) return PF_OK;
    if (K % 32 || N % 32 || A.ld % 8) return fail(PF_ERR_ARG, "tgemm: K/N must be multiples of 32");
    TmaGemmParams p{};
    p.M = (int)M; p.Cin = K; p.N = N; p.K = K; p.a_c0 = a_c0; p.groups = 1;
    fill_epi(p, w, o, 0);
    if (const char* d = getenv("PF_GEMM_DBG")) p.dbg = atoi(d);      // timing experiments only (gemm_tma.cuh: TmaGemmParams::dbg)
    TmaMaps maps{};
    const int bn = tma_pick_bn_gemm(M, N, K, e->sm_count);
    // CTA-pair kernel (256 x BN tiles, tcgen05.mma.cta_group::2: half the weight traffic per SM) when there are enough pair tiles
    const int pair_clusters = (e->use_pair && !getenv("PF_NO_PAIR")) ? gemm2_plan(e->device, (int)M, N, bn) : 0;
    const int kb = pair_clusters ? 32 : tma_pick_kb(bn, K, MODE_GEMM);
    const char* msg = nullptr;
    if (!msg) msg = map2d(&maps.a_hi, A.hi, A.ld, M, A.ld, 128, kb);
    if (!msg) msg = map2d(&maps.a_lo, A.lo, A.ld, M, A.ld, 128, kb);
    if (!msg) msg = map2d(&maps.b_hi, w.hi, K, N, K, pair_clusters ? bn / 2 : bn, kb);
    if (!msg) msg = map2d(&maps.b_lo, w.lo, K, N, K, pair_clusters ? bn / 2 : bn, kb);
    // epilogue tiles go through TMA as well: fp32 output, or (when there is no fp32 output) the split planes; residual
    if (o.C && o.S.hi) return fail(PF_ERR_ARG, "tgemm: fp32 and split outputs together are not supported in GEMM mode");
    if (o.res2 || o.bias_mode == 2) return fail(PF_ERR_ARG, "tgemm: second residual / border-class bias are halo-mode features");
    if (!msg && o.C) msg = map_tile32(&maps.c, o.C, M, o.ldc, true);
    if (!msg && o.S.hi) msg = map_tile32(&maps.s_hi, o.S.hi, M, o.S.ld, false);
    if (!msg && o.S.hi) msg = map_tile32(&maps.s_lo, o.S.lo, M, o.S.ld, false);
    if (!msg && o.res) msg = map_tile32(&maps.res, o.res, M, o.ldr, true);
    if (msg) return fail(PF_ERR_CUDA, "%s", msg);
    maps.a2_hi = maps.a_hi; maps.a2_lo = maps.a_lo;
    if (!o.C) maps.c = maps.a_hi;
    if (!o.S.hi) { maps.s_hi = maps.a_hi; maps.s_lo = maps.a_hi; }
    if (!o.res) maps.res = maps.a_hi;
    if (pair_clusters) return launch_pair(maps, p, bn, pair_clusters);
    return launch_tma(MODE_GEMM, maps, p);
  }
  static cudaError_t gemm_tma_pair_mode(const TmaMaps& maps, const TmaGemmParams& p, int bn, int nclusters, cudaStream_t st) {
    return gemm2_launch(maps, p, bn, nclusters, st);
  }
  int launch_pair(const TmaMaps& maps, const TmaGemmParams& p, int bn, int nclusters) {
    if (e->profile) {
      pf_engine::ProfRec r{};
      for (cudaEvent_t* ev : {&r.a, &r.b}) {
        if (e->ev_pool.empty()) { CU(cudaEventCreate(ev)); }
        else { *ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
      }
      r.flops = 2.0 * (double)p.M * (double)p.N * (double)p.K;
      r.cfg = 5;
      r.M = p.M; r.N = p.N; r.K = p.K; r.KH = 1; r.stride = 2; r.groups = 1; r.Cin = p.Cin;   // (stride 2 marks pair launches in the CSV)
      CU(cudaEventRecord(r.a, st));
      LAUNCHED(gemm_tma_pair_mode(maps, p, bn, nclusters, st));
      CU(cudaEventRecord(r.b, st));
      e->prof.push_back(r);
      return PF_OK;
    }
    LAUNCHED(gemm_tma_pair_mode(maps, p, bn, nclusters, st));
    return PF_OK;
  }
  // 3x3 / stride 1 / pad 1 convolution on split NHWC planes (optionally a second source for channels >= c_split)
  int thalo(const SplitT& A, int a_c0, int a_gc, const SplitT* A2, int c_split, int a2_c0, int B, int H, int W, int Cin, const GemmW& w, int N,
            int groups, int bias_gstride, const Epi& o, const PredTail* pred = nullptr) {
    if (dry) return PF_OK;
    if (Cin % 64 || N % 32 || (A2 && c_split % 64)) return fail(PF_ERR_ARG, "thalo: Cin must be a multiple of 64, N of 32");
    TmaGemmParams p{};
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.N = N; p.K = 9 * Cin; p.a_c0 = a_c0; p.a_gc = a_gc; p.groups = groups;
    p.c_split = A2 ? c_split : 0; p.a2_c0 = a2_c0;
    fill_epi(p, w, o, bias_gstride);
    if (const char* d = getenv("PF_HALO_DBG")) p.dbg = atoi(d);      // timing experiments only (gemm_tma.cuh: TmaGemmParams::dbg)
    TmaMaps maps{};
    const int bn = tma_pick_bn(N, MODE_HALO), kb = tma_pick_kb(bn, p.K, MODE_HALO);
    const char* msg = nullptr;
    if (!msg) msg = map_halo(&maps.a_hi, A.hi, B, H, W, A.ld);
    if (!msg) msg = map_halo(&maps.a_lo, A.lo, B, H, W, A.ld);
    if (A2) {
      if (!msg) msg = map_halo(&maps.a2_hi, A2->hi, B, H, W, A2->ld);
      if (!msg) msg = map_halo(&maps.a2_lo, A2->lo, B, H, W, A2->ld);
    } else { maps.a2_hi = maps.a_hi; maps.a2_lo = maps.a_lo; }
    if (!msg) msg = map2d(&maps.b_hi, w.hi, p.K, (long long)groups * N, p.K, bn, kb);
    if (!msg) msg = map2d(&maps.b_lo, w.lo, p.K, (long long)groups * N, p.K, bn, kb);
    maps.c = maps.a_hi; maps.s_hi = maps.a_hi; maps.s_lo = maps.a_hi; maps.res = maps.a_hi;   // halo mode: epilogue stores from registers
    if (msg) return fail(PF_ERR_CUDA, "%s", msg);
    return launch_tma(MODE_HALO, maps, p, pred);
  }
  // strided / patchifying convolution = patch gather on split planes + TMA GEMM
  int tconv_gather(const SplitT& A, int B, int H, int W, int Cin, int KH, int stride, int pad, const GemmW& w, int N, const Epi& o) {
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KH) / stride + 1;
    const long long M = (long long)B * OH * OW;
    const int K = KH * KH * Cin;
    const long long m = ar.mark();
    SplitT col = salloc(M, K);
    if (!dry) {
      if (Cin % 8) return fail(PF_ERR_ARG, "tconv_gather: Cin %% 8");
      LAUNCHED(launch_pdl(im2col_split_kernel, dim3(ew_grid(M * K / 8)), dim3(256), 0, st, A.hi, A.lo, A.ld, col.hi, col.lo, B, H, W, Cin, OH, OW, KH, stride, pad));
    }
    int r = tgemm(col, M, K, 0, w, N, o);
    ar.release(m);
    return r;
  }
  int ln_split(const float* x, const SplitT& y, long long rows, int C, const LnW& w, float eps, float* yf = nullptr) {
    if (dry) return PF_OK;
    LAUNCHED(layernorm_launch(x, yf, rows, C, w.w, w.b, eps, st, y));
    return PF_OK;
  }
  // attention core on tcgen05: q [n*N, C] and kv [n*100, 2C] split planes -> a split planes
  int attention_tc(const SplitT& q, const SplitT& kv, const SplitT& a, int nimg, int N, int C, int heads) {
    if (dry) return PF_OK;
    if (q.ld != C || kv.ld != 2 * C || a.ld != C || C != heads * kAtcD) return fail(PF_ERR_ARG, "attention_tc: layout");
    AtcMaps maps{};
    const char* msg = nullptr;
    if (!msg) msg = map2d(&maps.q_hi, q.hi, C, (long long)nimg * N, C, 128, 64);
    if (!msg) msg = map2d(&maps.q_lo, q.lo, C, (long long)nimg * N, C, 128, 64);
    if (!msg) msg = map2d(&maps.kv_hi, kv.hi, 2 * C, (long long)nimg * kAtcKeys, 2 * C, kAtcKeysPad, 64);
    if (!msg) msg = map2d(&maps.kv_lo, kv.lo, 2 * C, (long long)nimg * kAtcKeys, 2 * C, kAtcKeysPad, 64);
    if (msg) return fail(PF_ERR_CUDA, "%s", msg);
    LAUNCHED(attention_tc_launch(maps, a.hi, a.lo, nimg, N, C, heads, e->sm_count, st));
    return PF_OK;
  }
  // LayerNorm whose output is (also) written in patch order for a k = s = sr convolution on the R x R map (y may be empty)
  int ln_split_patch(const float* x, const SplitT& y, const SplitT& patch, long long rows, int C, const LnW& w, float eps, int R, int sr) {
    if (dry) return PF_OK;
    LAUNCHED(layernorm_launch(x, nullptr, rows, C, w.w, w.b, eps, st, y, patch, R, sr));
    return PF_OK;
  }
  int ln(const float* x, float* y, long long rows, int C, const LnW& w, float eps) {
    if (dry) return PF_OK;
    LAUNCHED(layernorm_launch(x, y, rows, C, w.w, w.b, eps, st));
    return PF_OK;
  }
};

// ----------------------------------------------------------------------------------------------- the forward graph
constexpr long long kTableSlabBytes = 8LL << 20;   // ~370 tables of a 2048-pixel axis; reset (after a stream sync) when full
static int get_table(pf_engine* e, int in_size, pf_engine::DevTable* out, cudaStream_t st) {
  auto it = e->tables.find(in_size);
  if (it == e->tables.end()) {
    ResampleTable t = make_resample_table(in_size, kNet);
    const long long nb = (long long)t.bounds.size() * sizeof(int), nc = (long long)t.coeffs.size() * sizeof(int);
    const long long need = ((nb + 255) & ~255LL) + ((nc + 255) & ~255LL);
    if (need > kTableSlabBytes) return fail(PF_ERR_ARG, "image axis of %d pixels is too long for the resize tables", in_size);
    if (e->table_off + need > kTableSlabBytes) {
      // slab full: earlier forwards on this stream may still read the old tables and the pinned mirror may still be the source
      // of an in-flight copy -> drain the stream once, then start over
      CU(cudaStreamSynchronize(st));
      e->tables.clear();
      e->table_off = 0;
    }
    pf_engine::DevTable d{};
    d.ksize = t.ksize;
    const long long o0 = e->table_off, o1 = o0 + ((nb + 255) & ~255LL);
    memcpy(e->table_host + o0, t.bounds.data(), nb);
    memcpy(e->table_host + o1, t.coeffs.data(), nc);
    d.bounds = (int*)(e->table_dev + o0);
    d.coeffs = (int*)(e->table_dev + o1);
    CU(cudaMemcpyAsync(e->table_dev + o0, e->table_host + o0, need, cudaMemcpyHostToDevice, st));
    e->table_off += need;
    it = e->tables.emplace(in_size, d).first;
  }
  *out = it->second;
  return PF_OK;
}

static int pre_rows_needed(int H) {  // input rows one block of kPreRows output rows may need
  const double scale = (double)H / kNet;
  const double support = scale < 1.0 ? 1.0 : scale;
  return (int)(kPreRows * scale) + 2 * (int)ceil(support) + 3;
}
constexpr int kPreMaxSmemRows = 200 * 1024 / (kNet * 3);

// ----------------------------------------------------------------------------------------------- shared graph sections
// uint8 HWC (any size) or pre-resized fp32 CHW -> x0 [n,320,320,4] fp32 normalised
static int fwd_preprocess(Fwd& F, const pf_batch* bt, float*& x0, PreImage*& d_pre, PostImage*& d_post) {
  pf_engine* e = F.e;
  const pf_model_desc& D = e->desc;
  const int n = F.n;
  const bool dry = F.dry;
  cudaStream_t st = F.st;
  Arena& ar = F.ar;
  // ---------------- pre-process: uint8 HWC (any size) -> [n,320,320,4] fp32 normalised -------------------
  x0 = ar.f((long long)n * kNet * kNet * 4);
  d_pre = (PreImage*)ar.alloc((long long)n * sizeof(PreImage));
  d_post = (PostImage*)ar.alloc((long long)n * sizeof(PostImage));
  if (!dry) {
    if (bt->images_u8) {
      std::vector<PreImage> pre(n);
      int max_h = 1;
      for (int i = 0; i < n; ++i) {
        const int H = bt->height[i], W = bt->width[i];
        if (H < 1 || W < 1) return fail(PF_ERR_ARG, "image %d has size %dx%d", i, H, W);
        pf_engine::DevTable tx, ty;
        TRY(get_table(e, W, &tx, st));
        TRY(get_table(e, H, &ty, st));
        if (ty.ksize + 1 > kPreMaxSmemRows) return fail(PF_ERR_ARG, "image %d is too tall (%d rows) for the resize kernel", i, H);
        pre[i] = PreImage{bt->image_offset[i], H, W, tx.ksize, ty.ksize, tx.bounds, tx.coeffs, ty.bounds, ty.coeffs};
        if (H > max_h) max_h = H;
      }
      CU(cudaMemcpyAsync(d_pre, pre.data(), n * sizeof(PreImage), cudaMemcpyHostToDevice, st));
      int rows = pre_rows_needed(max_h);
      if (rows > kPreMaxSmemRows) rows = kPreMaxSmemRows;
      const int smem = rows * kNet * 3;
      LAUNCHED((preprocess_kernel<<<dim3(kNet / kPreRows, n), kNet, smem, st>>>(bt->images_u8, d_pre, x0, D.pixel_mean[0], D.pixel_mean[1],
                                                                               D.pixel_mean[2], D.pixel_std[0], D.pixel_std[1], D.pixel_std[2], rows),
                cudaGetLastError()));
    } else {
      const long long total = (long long)n * kNet * kNet;
      LAUNCHED((normalize_chw_kernel<<<(unsigned)cdivl(total, 256), 256, 0, st>>>(bt->images_chw, x0, n, D.pixel_mean[0], D.pixel_mean[1],
                                                                                 D.pixel_mean[2], D.pixel_std[0], D.pixel_std[1], D.pixel_std[2]),
                cudaGetLastError()));
    }
  }
  return PF_OK;
}

// resample of the (decoded) 320x320 fields to the original sizes: one launch for all images of the batch
static int launch_postprocess(const float* vec, const float* lat, int n, const int32_t* height, const int32_t* width, const int64_t* g_off,
                              const int64_t* l_off, float* g_out, float* l_out, int lat_is_sin, PostImage* d_post, cudaStream_t st) {
  std::vector<PostImage> post(n);
  long long total = 0;
  int max_h = 1, max_wp = 4;
  for (int i = 0; i < n; ++i) {
    if (height[i] < 1 || width[i] < 1) return fail(PF_ERR_ARG, "image %d has size %dx%d", i, height[i], width[i]);
    post[i] = PostImage{height[i], width[i], g_off[i], l_off[i], total};
    total += (long long)height[i] * width[i];
    if (height[i] > max_h) max_h = height[i];
    const int wp = (width[i] + 3) / 4 * 4;
    if (wp <= kPostMaxW && wp > max_wp) max_wp = wp;    // (wider images take the table-less path of the kernel)
  }
  CU(cudaMemcpyAsync(d_post, post.data(), n * sizeof(PostImage), cudaMemcpyHostToDevice, st));
  const int smem = max_wp * 8;
  LAUNCHED((postprocess_kernel<<<dim3((unsigned)cdiv(max_h, kPostBand), (unsigned)n), kPostThreads, smem, st>>>(vec, lat, d_post, g_out, l_out, lat_is_sin),
            cudaGetLastError()));
  return PF_OK;
}

// prediction 1x1 convs (+ normalise / clamp) -> NCHW outputs, then argmax decode (classification) and resample to the
// original resolutions.  conv1_out: [n,320,320,64] fp32 (gravity head channels 0-31, latitude head 32-63).
static int fwd_tails_post(Fwd& F, const pf_batch* bt, const float* conv1_out, PostImage* d_post, bool pred_done = false) {
  pf_engine* e = F.e;
  const pf_model_desc& D = e->desc;
  const int n = F.n;
  const bool dry = F.dry;
  cudaStream_t st = F.st;
  Arena& ar = F.ar;
  const int HW = kNet * kNet;
  const bool cls_g = D.gravity_classes != 2, cls_l = D.latitude_classes != 1;
  const bool fused_decode = e->decode_only && (cls_g || cls_l);
  if (e->decode_only && cls_g != cls_l) return fail(PF_ERR_ARG, "decode_only needs both heads to be classification heads");
  const float* vec = dry ? nullptr : bt->pred_gravity;
  const float* lat = dry ? nullptr : bt->pred_latitude;
  if (e->debug) {
    // debug taps: the raw prediction-conv outputs before normalise / clamp (oracle taps g.raw / l.raw)
    float* rg = ar.f((long long)n * D.gravity_classes * HW);
    float* rl = ar.f((long long)n * D.latitude_classes * HW);
    if (!dry) {
      const unsigned grid = (unsigned)cdivl((long long)n * HW, 128);
      LAUNCHED((pred_tail_kernel<<<grid, 128, D.gravity_classes * 33 * 4, st>>>(conv1_out, 64, 0, e->pred_g_w, e->pred_g_b, rg, n, HW, D.gravity_classes, 0), cudaGetLastError()));
      LAUNCHED((pred_tail_kernel<<<grid, 128, D.latitude_classes * 33 * 4, st>>>(conv1_out, 64, 32, e->pred_l_w, e->pred_l_b, rl, n, HW, D.latitude_classes, 0), cudaGetLastError()));
    }
    TRY(F.tap("head.raw_g", rg, (long long)n * D.gravity_classes * HW));
    TRY(F.tap("head.raw_l", rl, (long long)n * D.latitude_classes * HW));
  }
  if (fused_decode) {
    // option "decode_only": 1x1 conv + argmax + bin decode in one kernel; pred_gravity / pred_latitude hold the decoded fields
    if (!dry) {
      const unsigned grid = ew_grid((long long)n * HW * 4);
      LAUNCHED((pred_argmax_decode_kernel<<<grid, 256, D.gravity_classes * 37 * 4, st>>>(conv1_out, 64, 0, e->pred_g_w, e->pred_g_b, bt->pred_gravity, n, HW,
                                                                                      D.gravity_classes, 1), cudaGetLastError()));
      LAUNCHED((pred_argmax_decode_kernel<<<grid, 256, D.latitude_classes * 37 * 4, st>>>(conv1_out, 64, 32, e->pred_l_w, e->pred_l_b, bt->pred_latitude, n, HW,
                                                                                       D.latitude_classes, 0), cudaGetLastError()));
    }
  } else {
    // prediction tails -> NCHW outputs (returned to the caller)
    if (!dry && !pred_done) {
      const unsigned grid = (unsigned)cdivl((long long)n * HW, 128);
      LAUNCHED((pred_tail_kernel<<<grid, 128, D.gravity_classes * 33 * 4, st>>>(conv1_out, 64, 0, e->pred_g_w, e->pred_g_b, bt->pred_gravity, n, HW,
                                                                             D.gravity_classes, D.gravity_classes == 2 ? 1 : 0), cudaGetLastError()));
      LAUNCHED((pred_tail_kernel<<<grid, 128, D.latitude_classes * 33 * 4, st>>>(conv1_out, 64, 32, e->pred_l_w, e->pred_l_b, bt->pred_latitude, n, HW,
                                                                              D.latitude_classes, D.latitude_classes == 1 ? 2 : 0), cudaGetLastError()));
    }
    if (cls_g) {
      float* dv = ar.f((long long)n * 2 * HW);
      if (!dry) LAUNCHED((argmax_decode_kernel<<<(unsigned)cdivl((long long)n * HW, 256), 256, 0, st>>>(bt->pred_gravity, dv, n, HW, D.gravity_classes, 1), cudaGetLastError()));
      vec = dv;
    }
    if (cls_l) {
      float* dl = ar.f((long long)n * HW);
      if (!dry) LAUNCHED((argmax_decode_kernel<<<(unsigned)cdivl((long long)n * HW, 256), 256, 0, st>>>(bt->pred_latitude, dl, n, HW, D.latitude_classes, 0), cudaGetLastError()));
      lat = dl;
    }
  }
  // ---------------- post-process to the original resolutions ------------------------------------------------
  if (!dry)
    TRY(launch_postprocess(vec, lat, n, bt->height, bt->width, bt->gravity_original_offset, bt->latitude_original_offset, bt->gravity_original,
                           bt->latitude_original, cls_l ? 0 : 1, d_post, st));
  return PF_OK;
}

// =============================================================================================== TMA forward graph
// Same network as run_forward, on the TMA -> tcgen05 engine: every GEMM input is a pre-split bf16 hi/lo tensor written by
// its producer (LayerNorm, attention, depthwise conv, upsample, stem, or the previous GEMM's epilogue).
static int run_forward_tma(Fwd& F, const pf_batch* bt) {
  pf_engine* e = F.e;
  const pf_model_desc& D = e->desc;
  const int n = F.n;
  const bool dry = F.dry;
  cudaStream_t st = F.st;
  Arena& ar = F.ar;
  using Epi = Fwd::Epi;

  float* x0; PreImage* d_pre; PostImage* d_post;
  {
    NvtxRange r_("pf:preprocess");
    TRY(fwd_preprocess(F, bt, x0, d_pre, d_post));
  }
  TRY(F.tap("pre", x0, (long long)n * kNet * kNet * 4));

  NvtxRange* sect = new NvtxRange("pf:ll_enc");
  struct SectGuard { NvtxRange*& p; ~SectGuard() { delete p; } } sect_guard{sect};
  auto section = [&](const char* name) { delete sect; sect = nullptr; sect = new NvtxRange(name); };
  SplitT cfeat[4];
  for (int s = 0; s < 4; ++s) cfeat[s] = F.salloc((long long)n * kMitRes[s] * kMitRes[s], kMitDims[s]);
  SplitT ll = F.salloc((long long)n * 160 * 160, 64);
  if (e->use_stem_tc) {   // conv7x7/2 (+ folded BN + ReLU) as patch gather + TMA GEMM (K = 147 padded to 160)
    const long long m = ar.mark();
    const long long M = (long long)n * 160 * 160;
    SplitT col = F.salloc(M, 160);
    if (!dry) LAUNCHED(launch_pdl(stem_gather_kernel, dim3(ew_grid(stem_gather_threads(n, 160, 160))), dim3(256), 0, st, x0, col.hi, col.lo, n, 160, 160, 2));
    Epi o; o.S = ll; o.act = 1;
    TRY(F.tgemm(col, M, 160, 0, e->llencg, 64, o));
    ar.release(m);
  } else if (!dry) {
    LAUNCHED((stem_conv_launch<7, 7, 2, 3, 64>(x0, 4, n, kNet, kNet, e->llenc_w, e->llenc_b, nullptr, 1, st, ll)));
  }
  TRY(F.tap_split("ll", ll, (long long)n * 160 * 160 * 64));

  // ---------------- MiT-B3 encoder ---------------------------------------------------------------------------
  for (int s = 0; s < 4; ++s) {
    { char nm[32]; snprintf(nm, sizeof nm, "pf:mit.stage%d", s + 1); section(nm); }
    const int C = kMitDims[s], R = kMitRes[s], N = R * R, heads = kMitHeads[s], sr = kMitSr[s];
    const long long rows = (long long)n * N;
    const long long m = ar.mark();
    float* x = ar.f(rows * C);
    float* tf = ar.f(rows * C);                      // patch-embed conv output (before its LayerNorm)
    SplitT t1 = F.salloc(rows, C);                   // LayerNorm output (GEMM input only)
    SplitT t1p;                                      // the same in patch order [n*100, sr*sr*C]: A operand of the spatial-reduction conv
    if (sr > 1) t1p = F.salloc((long long)n * 100, sr * sr * C);
    SplitT q = F.salloc(rows, C);                    // q and kv leave their GEMMs as split planes: the attention core's MMA operands
    SplitT a = F.salloc(rows, C);                    // attention output
    float* t2f = ar.f((long long)n * 100 * C);
    SplitT t2 = F.salloc((long long)n * 100, C);
    SplitT kv = F.salloc((long long)n * 100, 2 * C);
    const bool qkv_split = e->use_attn_mma && e->use_attn_split;
    float* qf = qkv_split ? nullptr : ar.f(rows * C);                        // (fp32 q / kv: CUDA-core attention kernel, or
    float* kvf = qkv_split ? nullptr : ar.f((long long)n * 100 * 2 * C);     //  option attn_split = 0)
    float* h1 = ar.f(rows * 4 * C);
    SplitT h2 = F.salloc(rows, 4 * C);
    if (s == 0 && e->use_stem_tc) {
      const long long mm = ar.mark();
      SplitT col = F.salloc(rows, 160);
      if (!dry) LAUNCHED(launch_pdl(stem_gather_kernel, dim3(ew_grid(stem_gather_threads(n, 80, 80))), dim3(256), 0, st, x0, col.hi, col.lo, n, 80, 80, 4));
      Epi o; o.C = tf; o.ldc = C;
      TRY(F.tgemm(col, rows, 160, 0, e->embed1g, 64, o));
      ar.release(mm);
    } else if (s == 0) {
      if (!dry) LAUNCHED((stem_conv_launch<7, 7, 4, 3, 64>(x0, 4, n, kNet, kNet, e->embed1_w, e->embed1_b, tf, 0, st)));
    } else {
      Epi o; o.C = tf; o.ldc = C;
      TRY(F.tconv_gather(cfeat[s - 1], n, kMitRes[s - 1], kMitRes[s - 1], kMitDims[s - 1], 3, 2, 1, e->embed[s], C, o));
    }
    TRY(F.ln(tf, x, rows, C, e->embed_ln[s], 1e-5f));
    TRY(F.tapf(x, rows * C, "mit.s%d.embed", s + 1));
    for (int i = 0; i < kMitDepths[s]; ++i) {
      const MitBlockW& b = e->blocks[s][i];
      if (sr > 1) TRY(F.ln_split_patch(x, t1, t1p, rows, C, b.ln1, 1e-6f, R, sr));     // + the sr conv's im2col matrix
      else TRY(F.ln_split(x, t1, rows, C, b.ln1, 1e-6f));
      Epi oq, okv;
      if (qkv_split) { oq.S = q; okv.S = kv; } else { oq.C = qf; oq.ldc = C; okv.C = kvf; okv.ldc = 2 * C; }
      // q and the spatial-reduction branch both depend on LayerNorm 1 only: fork the branch onto the engine's side stream (its
      // GEMMs have 50-75 tiles for 148 SMs; q's second, partial wave leaves SMs idle as well) and join before the attention core
      const bool fork = !dry && sr > 1 && e->use_fork && e->side && !e->kp.on && !e->profile && !sync_debug() && !e->debug;
      if (fork) {
        CU(cudaEventRecord(e->ev_fork, st));
        CU(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
        F.st = e->side;
      }
      if (sr > 1) {
        { Epi o; o.C = t2f; o.ldc = C; TRY(F.tgemm(t1p, (long long)n * 100, sr * sr * C, 0, b.sr, C, o)); }
        TRY(F.ln_split(t2f, t2, (long long)n * 100, C, b.srln, 1e-5f));
        TRY(F.tgemm(t2, (long long)n * 100, C, 0, b.kv, 2 * C, okv));
      }
      if (fork) {
        CU(cudaEventRecord(e->ev_join, e->side));
        F.st = st;
      }
      TRY(F.tgemm(t1, rows, C, 0, b.q, C, oq));
      if (fork) CU(cudaStreamWaitEvent(st, e->ev_join, 0));
      if (sr > 1) {
      } else {
        TRY(F.tgemm(t1, rows, C, 0, b.kv, 2 * C, okv));
      }
      if (qkv_split && e->use_attn_tc) {
        TRY(F.attention_tc(q, kv, a, n, N, C, heads));
      } else if (!dry) {
        if (qkv_split) LAUNCHED(attention_mma_launch(nullptr, nullptr, nullptr, n, N, C, heads, st, a, q, kv));
        else if (e->use_attn_mma) LAUNCHED(attention_mma_launch(qf, kvf, nullptr, n, N, C, heads, st, a));
        else LAUNCHED(attention_launch(qf, kvf, nullptr, n, N, C, heads, st, a));
      }
      { Epi o; o.C = x; o.ldc = C; o.res = x; o.ldr = C; TRY(F.tgemm(a, rows, C, 0, b.proj, C, o)); }
      TRY(F.tapf(x, rows * C, "mit.s%d.b%d.attn", s + 1, i));
      TRY(F.ln_split(x, t1, rows, C, b.ln2, 1e-6f));
      { Epi o; o.C = h1; o.ldc = 4 * C; TRY(F.tgemm(t1, rows, C, 0, b.fc1, 4 * C, o)); }
      if (!dry) LAUNCHED(launch_pdl(dwconv3x3_gelu_kernel, dim3(ew_grid((long long)n * ((R + 1) / 2) * ((R + PF_DW3_PX - 1) / PF_DW3_PX) * C)), dim3(256), 0, st, h1, nullptr, n, R, R, 4 * C, b.dw_w, b.dw_b, h2.hi, h2.lo));
      { Epi o; o.C = x; o.ldc = C; o.res = x; o.ldr = C; TRY(F.tgemm(h2, rows, 4 * C, 0, b.fc2, C, o)); }
      TRY(F.tapf(x, rows * C, "mit.s%d.b%d", s + 1, i));
    }
    TRY(F.ln_split(x, cfeat[s], rows, C, e->stage_norm[s], 1e-6f));
    { char nm[32]; snprintf(nm, sizeof nm, "mit.c%d", s + 1); TRY(F.tap_split(nm, cfeat[s], rows * C)); }
    ar.release(m);
  }

  // ---------------- decoder heads (group 0 = gravity, group 1 = latitude, side by side in the channel dimension) ----
  float* conv1_out = ar.f((long long)n * kNet * kNet * 64);
  bool fuse_pred = false;
  {
    const long long m = ar.mark();
    float* fused = nullptr;       // fp32 top-down feature of the previous level, upsampled to this level's resolution
    SplitT fused_s;               // level 1 only: the final fused feature at 160x160, split (input of conv_fuse_conv0)
    for (int lvl = 4; lvl >= 1; --lvl) {
      { char nm[32]; snprintf(nm, sizeof nm, "pf:heads.level%d", lvl); section(nm); }
      const int r = kMitRes[lvl - 1], Cin = kMitDims[lvl - 1];
      const long long px = (long long)n * r * r;
      float* t = ar.f(px * 512);
      SplitT rt = F.salloc(px, 512);        // relu(t)
      SplitT u = F.salloc(px, 512);         // rectified conv1 outputs
      float* v = ar.f(px * 512);
      SplitT rv = F.salloc(px, 512);        // relu(v)
      float* w2 = ar.f(px * 512);
      {   // composed linear_c{lvl} o linear_c{lvl}_proc (both heads: N = 512), border-class bias
        Epi o; o.C = t; o.ldc = 512; o.S = rt; o.split_relu = 1; o.bias_mode = 2;
        TRY(F.thalo(cfeat[lvl - 1], 0, 0, nullptr, 0, 0, n, r, r, Cin, e->proc[lvl - 1], 512, 1, 0, o));
        TRY(F.tapf(t, px * 512, "head.proc%d", lvl));
      }
      auto rcu = [&](const SplitT& A, const GemmW& w, Epi o) {
        o.c_gcoff = 256; o.s_gcoff = 256; o.r_gcoff = 256; o.r2_gcoff = 256;
        return F.thalo(A, 0, 256, nullptr, 0, 0, n, r, r, 256, w, 256, 2, 256, o);
      };
      const float* of = t;
      const SplitT* os = &rt;
      if (lvl < 4) {
        { Epi o; o.S = u; o.act = 1; TRY(rcu(rt, e->rcu[lvl - 1][0][0], o)); }
        { Epi o; o.C = v; o.ldc = 512; o.S = rv; o.split_relu = 1; o.res = t; o.ldr = 512; o.res_relu = 1; o.res2 = fused; o.ldr2 = 512;
          TRY(rcu(u, e->rcu[lvl - 1][0][1], o)); }
        of = v; os = &rv;
      }
      { Epi o; o.S = u; o.act = 1; TRY(rcu(*os, e->rcu[lvl - 1][1][0], o)); }
      { Epi o; o.C = w2; o.ldc = 512; o.res = of; o.ldr = 512; o.res_relu = 1; TRY(rcu(u, e->rcu[lvl - 1][1][1], o)); }
      if (lvl > 1) {
        float* up = ar.f(px * 4 * 512);
        if (!dry) LAUNCHED(launch_pdl(upsample2x_kernel, dim3(ew_grid(upsample2x_threads(n, r, r, 512))), dim3(256), 0, st, w2, 512, 0, up, 512, 0, n, r, r, 512, nullptr, nullptr));
        fused = up;
        TRY(F.tapf(up, px * 4 * 512, "head.fusion%d", lvl));
      } else {
        fused_s = F.salloc(px * 4, 512);
        if (!dry) LAUNCHED(launch_pdl(upsample2x_kernel, dim3(ew_grid(upsample2x_threads(n, r, r, 512))), dim3(256), 0, st, w2, 512, 0, nullptr, 512, 0, n, r, r, 512, fused_s.hi, fused_s.lo));
        TRY(F.tap_split("head.fusion1", fused_s, px * 4 * 512));
      }
    }
    // conv_fuse_conv0 on cat([fused, ll]) -> ReLU ; x2 ; conv_fuse_conv1 -> ReLU
    // regression heads: the 1x1 prediction conv + normalise / clamp run inside conv_fuse_conv1's epilogue (conv1's own output is
    // then only materialised for the debug taps); classification heads (73 / 180 logits) keep the separate tail kernel
    section("pf:heads.fuse_convs");
    fuse_pred = D.gravity_classes == 2 && D.latitude_classes == 1;
    const bool keep_conv1 = !fuse_pred || e->debug;   // (not `o.C != nullptr`: the sizing dry run has null pointers)
    PredTail pt[2] = {{e->pred_g_w, e->pred_g_b, dry ? nullptr : bt->pred_gravity, 2, 1}, {e->pred_l_w, e->pred_l_b, dry ? nullptr : bt->pred_latitude, 1, 2}};
    if (e->use_phase) {
      // x2 upsample folded into conv1's weights: conv1 runs on the 160x160 grid with N = 4 output phases x 32 (no upsampled
      // tensor); the two outermost output rows / columns, where the identity does not hold, are recomputed by conv1_ring_kernel
      SplitT c0s = F.salloc((long long)n * 160 * 160, 128);
      {
        Epi o; o.S = c0s; o.s_gcoff = 64; o.act = 1;
        TRY(F.thalo(fused_s, 0, 256, &ll, 256, 0, n, 160, 160, 320, e->conv0, 64, 2, 64, o));
        TRY(F.tap_split("head.conv0", c0s, (long long)n * 160 * 160 * 128));
      }
      Epi o; o.ldc = 64; o.c_gcoff = 32; o.act = 1; o.phase4 = 1;
      if (keep_conv1) o.C = conv1_out;
      TRY(F.thalo(c0s, 0, 64, nullptr, 0, 0, n, 160, 160, 64, e->conv1p, 128, 2, 128, o, fuse_pred ? pt : nullptr));
      if (!dry) {
        const dim3 grid((unsigned)cdiv(conv1_ring_count(kNet, kNet), kRingPx), (unsigned)n);
        LAUNCHED((conv1_ring_kernel<<<grid, 256, kRingSmem, st>>>(c0s.hi, c0s.lo, 160, 160, e->conv1f_w, e->conv1f_b, keep_conv1 ? conv1_out : nullptr,
                                                                 fuse_pred ? e->pred_g_w : nullptr, e->pred_g_b, bt->pred_gravity,
                                                                 fuse_pred ? e->pred_l_w : nullptr, e->pred_l_b, bt->pred_latitude), cudaGetLastError()));
      }
    } else {
      float* c0 = ar.f((long long)n * 160 * 160 * 128);
      {
        Epi o; o.C = c0; o.ldc = 128; o.c_gcoff = 64; o.act = 1;
        TRY(F.thalo(fused_s, 0, 256, &ll, 256, 0, n, 160, 160, 320, e->conv0, 64, 2, 64, o));
        TRY(F.tap("head.conv0", c0, (long long)n * 160 * 160 * 128));
      }
      SplitT c0u = F.salloc((long long)n * kNet * kNet, 128);
      if (!dry) LAUNCHED(launch_pdl(upsample2x_kernel, dim3(ew_grid(upsample2x_threads(n, 160, 160, 128))), dim3(256), 0, st, c0, 128, 0, nullptr, 128, 0, n, 160, 160, 128, c0u.hi, c0u.lo));
      Epi o; o.ldc = 64; o.c_gcoff = 32; o.act = 1;
      if (keep_conv1) o.C = conv1_out;
      TRY(F.thalo(c0u, 0, 64, nullptr, 0, 0, n, kNet, kNet, 64, e->conv1, 32, 2, 32, o, fuse_pred ? pt : nullptr));
    }
    if (keep_conv1) TRY(F.tap("head.conv1", conv1_out, (long long)n * kNet * kNet * 64));
    ar.release(m);
  }
  section("pf:tails_postprocess");
  TRY(fwd_tails_post(F, bt, conv1_out, d_post, fuse_pred));

  // ---------------- ParamNet (ConvNeXt-T on the predicted fields) -------------------------------------------
  if (D.param_net != PF_PARAM_NONE) {
    section("pf:paramnet");
    if (D.gravity_classes != 2 || D.latitude_classes != 1) return fail(PF_ERR_ARG, "ParamNet needs regression heads");
    const int S = D.param_net == PF_PARAM_CENTERED ? kNet : D.param_input_size;
    float* pin = ar.f((long long)n * S * S * 4);
    if (!dry) LAUNCHED((pack_fields_kernel<<<(unsigned)cdivl((long long)n * S * S, 256), 256, 0, st>>>(bt->pred_gravity, bt->pred_latitude, pin, n, S), cudaGetLastError()));
    int r = S / 4;
    float* x = ar.f((long long)n * r * r * 96);
    if (!dry) LAUNCHED((stem_conv_launch<4, 4, 4, 0, 96>(pin, 4, n, S, S, e->pn_stem_w, e->pn_stem_b, x, 0, st)));
    TRY(F.ln(x, x, (long long)n * r * r, 96, e->pn_stem_ln, 1e-6f));
    for (int s = 0; s < 4; ++s) {
      const int C = kCnxDims[s];
      if (s > 0) {
        const int r2 = r / 2;
        SplitT y = F.salloc((long long)n * r2 * r2, 4 * kCnxDims[s - 1]);      // LayerNorm output written directly as the 2x2/2 conv's im2col matrix
        TRY(F.ln_split_patch(x, SplitT(), y, (long long)n * r * r, kCnxDims[s - 1], e->pn_ds_ln[s], 1e-6f, r, 2));
        float* xn = ar.f((long long)n * r2 * r2 * C);
        Epi o; o.C = xn; o.ldc = C;
        TRY(F.tgemm(y, (long long)n * r2 * r2, 4 * kCnxDims[s - 1], 0, e->pn_ds[s], C, o));
        x = xn; r = r2;
      }
      const long long rows = (long long)n * r * r;
      float* yf = ar.f(rows * C);
      SplitT y = F.salloc(rows, C);
      SplitT h = F.salloc(rows, 4 * C);
      for (int j = 0; j < kCnxDepths[s]; ++j) {
        const CnxBlockW& b = e->pn_blocks[s][j];
        if (e->use_dwln) {   // depthwise 7x7 + LayerNorm in one kernel, straight to the split planes pwconv1 loads
          if (!dry) LAUNCHED(dwconv7x7_ln_launch(x, n, r, r, C, b.dw_w, b.dw_b, b.ln.w, b.ln.b, 1e-6f, y, st));
        } else {
          if (!dry) LAUNCHED(launch_pdl(dwconv7x7_kernel, dim3(ew_grid((long long)n * ((r + 1) / 2) * ((r + PF_DW7_PX - 1) / PF_DW7_PX) * (C / 4))), dim3(256), 0, st, x, yf, n, r, r, C, b.dw_w, b.dw_b));
          TRY(F.ln_split(yf, y, rows, C, b.ln, 1e-6f));
        }
        { Epi o; o.S = h; o.act = 2; TRY(F.tgemm(y, rows, C, 0, b.pw1, 4 * C, o)); }
        { Epi o; o.C = x; o.ldc = C; o.res = x; o.ldr = C; o.gamma = b.gamma; TRY(F.tgemm(h, rows, 4 * C, 0, b.pw2, C, o)); }
      }
      TRY(F.tapf(x, rows * C, "cnx.s%d", s));
    }
    if (!dry) {
      if (!bt->params) return fail(PF_ERR_ARG, "params output is NULL");
      LAUNCHED((param_tail_kernel<<<n, 256, 0, st>>>(x, r * r, e->pn_norm.w, e->pn_norm.b, e->pn_head_w, e->pn_head_b, bt->params, D.param_net), cudaGetLastError()));
    }
  }
  return PF_OK;
}

// ----------------------------------------------------------------------------------------------- C ABI
extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }
const char* pf_last_error(void) { return g_err.c_str(); }
int64_t pf_kernel_launch_count(void) { return g_launches.load(); }

// The opt-in for more than 48 KB of dynamic shared memory is a per-device attribute of each kernel: set for every kernel of the
// library on every device an engine (or an operator entry point) uses, once per device and thread-safe.
static int configure_device(int device) {
  static std::mutex mu;
  static std::vector<char> done;
  std::lock_guard<std::mutex> lock(mu);
  if (device < (int)done.size() && done[device]) return PF_OK;
  CU(gemm_tma_configure_device());
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    CU(gemm2_configure_device(device, prop.multiProcessorCount));
  }
  CU(attention_mma_configure_device());
  CU(attention_tc_configure_device());
  CU(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  CU(cudaFuncSetAttribute(conv1_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));
  CU(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreMaxSmemRows * kNet * 3));
  if (device >= (int)done.size()) done.resize(device + 1, 0);
  done[device] = 1;
  return PF_OK;
}
static int configure_current_device() {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  return configure_device(dev);
}

int pf_create(int device, const pf_model_desc* desc, pf_handle* out) {
  if (!desc || !out) return fail(PF_ERR_ARG, "pf_create: null argument");
  if (!((desc->gravity_classes == 2 || desc->gravity_classes == 73) && (desc->latitude_classes == 1 || desc->latitude_classes == 180)))
    return fail(PF_ERR_ARG, "pf_create: unsupported head widths %d/%d", desc->gravity_classes, desc->latitude_classes);
  if (desc->param_net < 0 || desc->param_net > 2) return fail(PF_ERR_ARG, "pf_create: bad param_net");
  if (desc->param_net == PF_PARAM_UNCENTERED && (desc->param_input_size < 32 || desc->param_input_size > kNet || desc->param_input_size % 32))
    return fail(PF_ERR_ARG, "pf_create: param_input_size must be a multiple of 32 in [32, 320]");
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(PF_ERR_CUDA, "pf_create: no CUDA device %d (found %d)", device, count);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(PF_ERR_CUDA, "pf_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  TRY(configure_device(device));
  pf_engine* e = new pf_engine();
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->desc = *desc;
  if (cudaMalloc(&e->table_dev, kTableSlabBytes) != cudaSuccess || cudaMallocHost(&e->table_host, kTableSlabBytes) != cudaSuccess) {
    const int r = fail(PF_ERR_CUDA, "pf_create: resize-table slab: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(e->table_dev);
    delete e;
    return r;
  }
  if (cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    const int r = fail(PF_ERR_CUDA, "pf_create: side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
    pf_destroy(e);
    return r;
  }
  *out = e;
  return PF_OK;
}

int pf_destroy(pf_handle h) {
  if (!h) return PF_OK;
  cudaSetDevice(h->device);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFree(h->table_dev);
  cudaFreeHost(h->table_host);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto ev : h->ev_pool) cudaEventDestroy(ev);
  for (auto ev : h->kp.pool) cudaEventDestroy(ev);
  delete h;
  return PF_OK;
}

int pf_set_weight(pf_handle h, const char* name, const void* dev_ptr, int64_t numel, int dtype) {
  if (!h || !name || !dev_ptr) return fail(PF_ERR_ARG, "pf_set_weight: null argument");
  if (((uintptr_t)dev_ptr & 15) != 0) return fail(PF_ERR_ARG, "pf_set_weight: '%s' is not 16-byte aligned", name);
  h->weights[name] = WeightRef{dev_ptr, numel, dtype};
  h->finalized = false;
  return PF_OK;
}

int pf_finalize(pf_handle h) {
  if (!h) return fail(PF_ERR_ARG, "pf_finalize: null handle");
  TRY(resolve_weights(h));
  h->finalized = true;
  return PF_OK;
}

int64_t pf_workspace_bytes(pf_handle h, int n, int max_h) {
  if (!h || n < 1) return fail(PF_ERR_ARG, "pf_workspace_bytes: bad argument");
  Fwd F{h, Arena{}, nullptr, true, n};
  F.ar.dry = true;
  F.ar.keep = h->debug;
  (void)max_h;
  int r = run_forward_tma(F, nullptr);
  if (r != PF_OK) return r;
  return F.ar.peak + 4096;
}

int pf_forward(pf_handle h, const pf_batch* bt, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !bt || !workspace) return fail(PF_ERR_ARG, "pf_forward: null argument");
  if (!h->finalized) return fail(PF_ERR_WEIGHT, "pf_forward: pf_finalize has not succeeded");
  if (bt->n < 1) return fail(PF_ERR_ARG, "pf_forward: empty batch");
  if ((bt->images_u8 != nullptr) == (bt->images_chw != nullptr)) return fail(PF_ERR_ARG, "pf_forward: exactly one of images_u8 / images_chw");
  if (bt->images_u8 && !bt->image_offset) return fail(PF_ERR_ARG, "pf_forward: image_offset is NULL");
  if (!bt->height || !bt->width || !bt->pred_gravity || !bt->pred_latitude || !bt->gravity_original || !bt->latitude_original ||
      !bt->gravity_original_offset || !bt->latitude_original_offset)
    return fail(PF_ERR_ARG, "pf_forward: null input/output pointer");
  CU(cudaSetDevice(h->device));
  Fwd F{h, Arena{}, (cudaStream_t)stream, false, bt->n};
  F.ar.base = (char*)workspace;
  F.ar.cap = workspace_bytes;
  F.ar.keep = h->debug;
  {  // capacity check with a dry run (cheap: no launches)
    Fwd T{h, Arena{}, nullptr, true, bt->n};
    T.ar.dry = true; T.ar.keep = h->debug;
    TRY(run_forward_tma(T, nullptr));
    if (T.ar.peak > workspace_bytes) return fail(PF_ERR_WORKSPACE, "pf_forward: workspace %lld B < required %lld B", (long long)workspace_bytes, T.ar.peak);
  }
  if (((uintptr_t)workspace & 255) != 0) return fail(PF_ERR_ARG, "pf_forward: workspace must be 256-byte aligned");
  h->taps.clear();
  h->kp.st = (cudaStream_t)stream;
  tl_kp = &h->kp;
  pdl_enabled() = h->use_pdl && !h->kp.on && !h->profile && !sync_debug();   // (event records between launches defeat it anyway)
  const int r = run_forward_tma(F, bt);
  pdl_enabled() = false;
  tl_kp = nullptr;
  return r;
}

int pf_profile_enable(pf_handle h, int on) {
  if (!h) return fail(PF_ERR_ARG, "null handle");
  h->profile = on != 0;
  // `on` > 1: pre-create the CUDA events for that many GEMM launches now, so none is created inside a timed region
  while (on > 1 && (long long)h->ev_pool.size() < 2LL * on) {
    cudaEvent_t ev;
    CU(cudaEventCreate(&ev));
    h->ev_pool.push_back(ev);
  }
  return PF_OK;
}
int pf_profile_kernels_enable(pf_handle h, int max_launches) {
  if (!h) return fail(PF_ERR_ARG, "null handle");
  CU(cudaSetDevice(h->device));
  KernelProf& kp = h->kp;
  kp.on = max_launches > 0;
  kp.used = 0;
  kp.recs.clear();
  while ((long long)kp.pool.size() < 2LL * max_launches) {
    cudaEvent_t ev;
    CU(cudaEventCreate(&ev));
    kp.pool.push_back(ev);
  }
  return PF_OK;
}
// text table "kernel,launches,ms\n" aggregated over the launches recorded since pf_profile_kernels_enable; the caller must have
// synchronised the stream.  Returns the number of bytes written (excluding the terminating NUL) or a negative status.
int pf_profile_kernels_read(pf_handle h, char* buf, int cap) {
  if (!h || !buf || cap < 1) return fail(PF_ERR_ARG, "pf_profile_kernels_read: bad argument");
  std::map<std::string, std::pair<int, double>> agg;
  std::vector<std::string> order;
  for (const auto& r : h->kp.recs) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.a, r.b));
    const char* c = r.expr;
    while (*c == '(' || *c == ' ') ++c;
    const char* e = c;
    while (*e && (isalnum((unsigned char)*e) || *e == '_')) ++e;
    std::string name(c, e);
    if (name == "launch_pdl" && *e == '(') {   // launch_pdl(kernel, grid, ...): the kernel is the first argument
      c = e + 1;
      e = c;
      while (*e && (isalnum((unsigned char)*e) || *e == '_')) ++e;
      name.assign(c, e);
    }
    // template arguments of direct kernel launches distinguish the variants (e.g. stem_conv_launch<7, 7, 2, 3, 64>)
    if (*e == '<' && e[1] != '<') { const char* t = strchr(e, '>'); if (t) name.append(e, t + 1); }
    auto it = agg.find(name);
    if (it == agg.end()) { order.push_back(name); it = agg.emplace(name, std::make_pair(0, 0.0)).first; }
    it->second.first += 1;
    it->second.second += ms;
  }
  std::string out = "kernel,launches,ms\n";
  for (const auto& n : order) {
    char line[256];
    snprintf(line, sizeof line, "%s,%d,%.4f\n", n.c_str(), agg[n].first, agg[n].second);
    out += line;
  }
  if ((int)out.size() + 1 > cap) return fail(PF_ERR_ARG, "pf_profile_kernels_read: buffer too small (%d needed)", (int)out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  h->kp.used = 0;
  h->kp.recs.clear();
  return (int)out.size();
}
int pf_set_option(pf_handle h, const char* name, int value) {
  if (!h || !name) return fail(PF_ERR_ARG, "pf_set_option: null argument");
  if (!strcmp(name, "attn_mma")) { h->use_attn_mma = value != 0; return PF_OK; }
  if (!strcmp(name, "stem_tc")) { h->use_stem_tc = value != 0; return PF_OK; }
  if (!strcmp(name, "phase_conv1")) { h->use_phase = value != 0; return PF_OK; }
  if (!strcmp(name, "attn_split")) { h->use_attn_split = value != 0; return PF_OK; }
  if (!strcmp(name, "decode_only")) { h->decode_only = value != 0; return PF_OK; }
  if (!strcmp(name, "pdl")) { h->use_pdl = value != 0; return PF_OK; }
  if (!strcmp(name, "dw_ln")) { h->use_dwln = value != 0; return PF_OK; }
  if (!strcmp(name, "pair")) { h->use_pair = value != 0; return PF_OK; }
  if (!strcmp(name, "fork")) { h->use_fork = value != 0; return PF_OK; }
  if (!strcmp(name, "attn_tc")) { h->use_attn_tc = value != 0; return PF_OK; }
  return fail(PF_ERR_ARG, "pf_set_option: unknown option '%s'", name);
}
// out[cfg*3 + {0,1,2}] = {milliseconds, algorithmic FLOPs, launches} per GEMM engine configuration (7 configs),
// accumulated since the last read; the caller must have synchronised the stream.
int pf_profile_read(pf_handle h, double* out9) {
  if (!h || !out9) return fail(PF_ERR_ARG, "pf_profile_read: null argument");
  for (int i = 0; i < 21; ++i) out9[i] = 0.0;
  FILE* csv = nullptr;
  if (const char* path = getenv("PF_PROFILE_CSV")) {   // optional per-launch dump (profiles/)
    csv = fopen(path, "w");
    if (csv) fprintf(csv, "engine_cfg,M,N,K,Cin,KH,stride,groups,ms,algorithmic_tflops\n");
  }
  for (auto& r : h->prof) {
    float ms = 0.f;
    CU(cudaEventSynchronize(r.b));
    CU(cudaEventElapsedTime(&ms, r.a, r.b));
    if (csv) fprintf(csv, "%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.1f\n", r.cfg, r.M, r.N, r.K, r.Cin, r.KH, r.stride, r.groups, ms, r.flops / (ms * 1e9));
    out9[r.cfg * 3 + 0] += ms;
    out9[r.cfg * 3 + 1] += r.flops;
    out9[r.cfg * 3 + 2] += 1.0;
    h->ev_pool.push_back(r.a);
    h->ev_pool.push_back(r.b);
  }
  if (csv) fclose(csv);
  h->prof.clear();
  return PF_OK;
}

int pf_debug_enable(pf_handle h, int on) {
  if (!h) return fail(PF_ERR_ARG, "null handle");
  h->debug = on != 0;
  h->taps.clear();
  return PF_OK;
}
int pf_debug_count(pf_handle h) { return h ? (int)h->taps.size() : 0; }
const char* pf_debug_name(pf_handle h, int i) { return (h && i >= 0 && i < (int)h->taps.size()) ? h->taps[i].first.c_str() : ""; }
int64_t pf_debug_numel(pf_handle h, const char* name) {
  if (!h || !name) return -1;
  for (auto& t : h->taps) if (t.first == name) return t.second.second;
  return -1;
}
int pf_debug_copy(pf_handle h, const char* name, float* dst, int64_t numel, void* stream) {
  if (!h || !name || !dst) return fail(PF_ERR_ARG, "pf_debug_copy: null argument");
  for (auto& t : h->taps)
    if (t.first == name) {
      if (numel != t.second.second) return fail(PF_ERR_ARG, "pf_debug_copy: '%s' has %lld elements", name, t.second.second);
      CU(cudaMemcpyAsync(dst, t.second.first, numel * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
      return PF_OK;
    }
  return fail(PF_ERR_ARG, "pf_debug_copy: no tap '%s'", name);
}

// ---- multi-GPU gather (NCCL point-to-point) ---------------------------------------------------
#define NCCL_TRY(expr)                                                                                              \
  do {                                                                                                              \
    int r__ = (expr);                                                                                               \
    if (r__ != kNcclSuccess) return fail(PF_ERR_CUDA, "%s: %s", #expr, api.GetErrorString ? api.GetErrorString(r__) : "NCCL error"); \
  } while (0)
int pf_comm_unique_id(void* id128) {
  if (!id128) return fail(PF_ERR_ARG, "pf_comm_unique_id: null argument");
  const NcclApi& api = nccl_api();
  if (api.error) return fail(PF_ERR_CUDA, "%s", api.error);
  static_assert(sizeof(NcclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NCCL_TRY(api.GetUniqueId((NcclUniqueId*)id128));
  return PF_OK;
}
int pf_comm_create(int device, int rank, int nranks, const void* id128, pf_comm_handle* out) {
  if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(PF_ERR_ARG, "pf_comm_create: bad argument");
  const NcclApi& api = nccl_api();
  if (api.error) return fail(PF_ERR_CUDA, "%s", api.error);
  CU(cudaSetDevice(device));
  NcclUniqueId id;
  memcpy(&id, id128, sizeof id);
  pf_comm* c = new pf_comm();
  c->device = device; c->rank = rank; c->nranks = nranks;
  const int r = api.CommInitRank(&c->comm, nranks, id, rank);
  if (r != kNcclSuccess) { delete c; return fail(PF_ERR_CUDA, "ncclCommInitRank: %s", api.GetErrorString(r)); }
  *out = c;
  return PF_OK;
}
int pf_comm_destroy(pf_comm_handle c) {
  if (!c) return PF_OK;
  const NcclApi& api = nccl_api();
  cudaSetDevice(c->device);
  if (c->comm && api.CommDestroy) api.CommDestroy(c->comm);
  delete c;
  return PF_OK;
}
int pf_gather(pf_comm_handle c, int root, int count, void* const* dev_ptrs, const int64_t* bytes, const int32_t* peer, void* stream) {
  if (!c || count < 0 || root < 0 || root >= c->nranks || (count > 0 && (!dev_ptrs || !bytes))) return fail(PF_ERR_ARG, "pf_gather: bad argument");
  if (c->rank == root && count > 0 && !peer) return fail(PF_ERR_ARG, "pf_gather: the root needs the source rank of every segment");
  const NcclApi& api = nccl_api();
  CU(cudaSetDevice(c->device));
  if (count == 0) return PF_OK;
  NCCL_TRY(api.GroupStart());
  for (int i = 0; i < count; ++i) {
    int r;
    if (c->rank == root) {
      if (peer[i] < 0 || peer[i] >= c->nranks || peer[i] == root) { api.GroupEnd(); return fail(PF_ERR_ARG, "pf_gather: segment %d comes from rank %d", i, peer[i]); }
      r = api.Recv(dev_ptrs[i], (size_t)bytes[i], kNcclUint8, peer[i], c->comm, (cudaStream_t)stream);
    } else {
      r = api.Send(dev_ptrs[i], (size_t)bytes[i], kNcclUint8, root, c->comm, (cudaStream_t)stream);
    }
    if (r != kNcclSuccess) { api.GroupEnd(); return fail(PF_ERR_CUDA, "ncclSend/Recv: %s", api.GetErrorString(r)); }
  }
  NCCL_TRY(api.GroupEnd());
  return PF_OK;
}

// ---- decode front-end (nvJPEG) ------------------------------------------------------------------
int pf_jpeg_create(int device, int max_threads, pf_jpeg_handle* out) {
  if (!out) return fail(PF_ERR_ARG, "pf_jpeg_create: null argument");
  const NvjpegApi& api = nvjpeg_api();
  if (api.error) return fail(PF_ERR_CUDA, "%s", api.error);
  CU(cudaSetDevice(device));
  pf_jpeg* j = new pf_jpeg();
  j->device = device;
  if (api.CreateSimple(&j->handle) != NVJPEG_STATUS_SUCCESS) { delete j; return fail(PF_ERR_CUDA, "nvjpegCreateSimple failed"); }
  int nt = max_threads > 0 ? max_threads : (int)std::thread::hardware_concurrency() / 2;
  nt = nt < 1 ? 1 : (nt > 32 ? 32 : nt);
  j->workers.resize(nt);
  bool ok = cudaEventCreateWithFlags(&j->start, cudaEventDisableTiming) == cudaSuccess;
  for (auto& w : j->workers) {
    ok = ok && api.StateCreate(j->handle, &w.state) == NVJPEG_STATUS_SUCCESS;
    ok = ok && cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) { pf_jpeg_destroy(j); return fail(PF_ERR_CUDA, "pf_jpeg_create: decoder state / stream creation failed"); }
  *out = j;
  return PF_OK;
}
int pf_jpeg_destroy(pf_jpeg_handle j) {
  if (!j) return PF_OK;
  const NvjpegApi& api = nvjpeg_api();
  cudaSetDevice(j->device);
  for (auto& w : j->workers) {
    if (w.stream) cudaStreamSynchronize(w.stream);
    if (w.state) api.StateDestroy(w.state);
    if (w.stream) cudaStreamDestroy(w.stream);
    if (w.done) cudaEventDestroy(w.done);
  }
  if (j->start) cudaEventDestroy(j->start);
  if (j->handle) api.Destroy(j->handle);
  delete j;
  return PF_OK;
}
int pf_jpeg_info(pf_jpeg_handle j, const uint8_t* data, int64_t length, int32_t* height, int32_t* width) {
  if (!j || !data || length < 4 || !height || !width) return fail(PF_ERR_ARG, "pf_jpeg_info: bad argument");
  const NvjpegApi& api = nvjpeg_api();
  int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  if (api.GetImageInfo(j->handle, data, (size_t)length, &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) return fail(PF_ERR_ARG, "pf_jpeg_info: not a decodable JPEG stream");
  *height = hs[0]; *width = ws[0];
  return PF_OK;
}
int pf_jpeg_decode_batch(pf_jpeg_handle j, int n, const uint8_t* const* data, const int64_t* length, const int32_t* height, const int32_t* width,
                         uint8_t* blob, const int64_t* offset, void* stream) {
  if (!j || n < 1 || !data || !length || !height || !width || !blob || !offset) return fail(PF_ERR_ARG, "pf_jpeg_decode_batch: bad argument");
  const NvjpegApi& api = nvjpeg_api();
  CU(cudaSetDevice(j->device));
  cudaStream_t st = (cudaStream_t)stream;
  // the workers' streams start after everything already queued on the caller's stream (the blob may be in use by an earlier forward)
  CU(cudaEventRecord(j->start, st));
  const int nt = (int)j->workers.size() < n ? (int)j->workers.size() : n;
  std::atomic<int> next{0}, failed{-1};
  auto work = [&](int t) {
    cudaSetDevice(j->device);
    pf_jpeg::Worker& w = j->workers[t];
    cudaStreamWaitEvent(w.stream, j->start, 0);
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      nvjpegImage_t dst{};
      dst.channel[0] = blob + offset[i];
      dst.pitch[0] = (size_t)width[i] * 3;
      if (api.Decode(j->handle, w.state, data[i], (size_t)length[i], NVJPEG_OUTPUT_BGRI, &dst, w.stream) != NVJPEG_STATUS_SUCCESS) failed.store(i);
    }
    cudaEventRecord(w.done, w.stream);
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < nt; ++t) threads.emplace_back(work, t);
  work(0);
  for (auto& th : threads) th.join();
  for (int t = 0; t < nt; ++t) CU(cudaStreamWaitEvent(st, j->workers[t].done, 0));
  if (failed.load() >= 0) return fail(PF_ERR_ARG, "pf_jpeg_decode_batch: image %d could not be decoded", failed.load());
  return PF_OK;
}

// ---- single-operator entry points ------------------------------------------------------------------------
int pf_op_conv_gemm(const float* x, int B, int H, int W, int Cin, const void* whi, const void* wlo, const float* bias, int N, int KH, int KW,
                    int stride, int pad, int in_relu, int act, const float* res, int res_relu, float* y, void* stream) {
  // split the input the way a producer kernel would, then run the same helpers the forward graph uses
  if (!x || !whi || !wlo || !y) return fail(PF_ERR_ARG, "pf_op_conv_gemm: null argument");
  if (KH != KW) return fail(PF_ERR_ARG, "pf_op_conv_gemm: square filters only");
  TRY(configure_current_device());
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  pf_engine tmp;
  tmp.device = dev;
  tmp.sm_count = prop.multiProcessorCount;
  const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad - KW) / stride + 1;
  const long long nx = (long long)B * H * W * Cin;
  const long long colb = (long long)B * OH * OW * KH * KW * Cin * 4 + (1 << 20);
  char* scratch = nullptr;
  CU(cudaMalloc(&scratch, nx * 4 + colb + 4096));
  Fwd F{&tmp, Arena{}, st, false, B};
  F.ar.base = scratch; F.ar.cap = nx * 4 + colb + 4096;
  SplitT A = F.salloc((long long)B * H * W, Cin);
  int r = PF_OK;
  {
    cudaError_t le = (split_kernel<<<(unsigned)cdivl(nx, 256), 256, 0, st>>>(x, A.hi, A.lo, nx, in_relu), cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (le != cudaSuccess) r = fail(PF_ERR_CUDA, "split_kernel: %s", cudaGetErrorString(le));
  }
  GemmW w{(const __nv_bfloat16*)whi, (const __nv_bfloat16*)wlo, bias};
  Fwd::Epi o;
  o.C = y; o.ldc = N; o.act = act; o.res = res; o.ldr = N; o.res_relu = res_relu;
  if (r == PF_OK) {
    if (KH == 3 && stride == 1 && pad == 1 && Cin % 64 == 0) r = F.thalo(A, 0, 0, nullptr, 0, 0, B, H, W, Cin, w, N, 1, 0, o);
    else if (KH == 1 && stride == 1 && pad == 0) r = F.tgemm(A, (long long)B * H * W, Cin, 0, w, N, o);
    else r = F.tconv_gather(A, B, H, W, Cin, KH, stride, pad, w, N, o);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(scratch);
  if (r != PF_OK) return r;
  if (se != cudaSuccess) return fail(PF_ERR_CUDA, "pf_op_conv_gemm: %s", cudaGetErrorString(se));
  return PF_OK;
}
int pf_camera_fields(int device, const pf_camera* cams, int n, float* up, float* lat, void* stream) {
  if (!cams || n < 1 || (!up && !lat)) return fail(PF_ERR_ARG, "pf_camera_fields: bad argument");
  CU(cudaSetDevice(device));
  for (int i0 = 0; i0 < n; i0 += kCamChunk) {
    const int m = n - i0 < kCamChunk ? n - i0 : kCamChunk;
    CamBatch b{};
    long long max_q = 1;
    for (int i = 0; i < m; ++i) {
      const pf_camera& c = cams[i0 + i];
      if (c.height < 1 || c.width < 1 || !(c.focal_rel != 0.0)) return fail(PF_ERR_ARG, "pf_camera_fields: image %d: size %dx%d, focal %g", i0 + i, c.height, c.width, c.focal_rel);
      if ((c.up_offset & 1) != 0) return fail(PF_ERR_ARG, "pf_camera_fields: up_offset must be even (8-byte stores)");
      CamImage& o = b.im[i];
      o.H = c.height; o.W = c.width;
      o.f = c.focal_rel * c.height;
      o.cx = (c.cx_rel + 0.5) * c.width; o.cy = (c.cy_rel + 0.5) * c.height;
      o.sr = sin(c.roll); o.cr = cos(c.roll); o.se = sin(c.elevation); o.ce = cos(c.elevation);
      o.sgn = c.elevation > 0 ? 1.0 : (c.elevation < 0 ? -1.0 : 0.0);
      o.up_off = c.up_offset; o.lat_off = c.lat_offset;
      const long long qd = (long long)c.height * ((c.width + 3) / 4);
      if (qd > max_q) max_q = qd;
    }
    const dim3 grid((unsigned)cdivl(max_q, 256), (unsigned)m);
    LAUNCHED((camera_fields_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, up, lat), cudaGetLastError()));
  }
  return PF_OK;
}

static int upload_table(const ResampleTable& t, int** bounds, int** coeffs) {
  CU(cudaMalloc(bounds, t.bounds.size() * 4));
  CU(cudaMalloc(coeffs, t.coeffs.size() * 4));
  CU(cudaMemcpy(*bounds, t.bounds.data(), t.bounds.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(*coeffs, t.coeffs.data(), t.coeffs.size() * 4, cudaMemcpyHostToDevice));
  return PF_OK;
}
int pf_op_resize_u8(const uint8_t* img, int H, int W, int new_h, int new_w, uint8_t* out, void* stream) {
  if (!img || !out || H < 1 || W < 1 || new_h < 1 || new_w < 1) return fail(PF_ERR_ARG, "pf_op_resize_u8: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (H == new_h && W == new_w) {   // Pillow returns a copy when the size does not change
    CU(cudaMemcpyAsync(out, img, (size_t)H * W * 3, cudaMemcpyDeviceToDevice, st));
    return PF_OK;
  }
  ResampleTable tx = make_resample_table(W, new_w), ty = make_resample_table(H, new_h);
  int *bx = nullptr, *cx = nullptr, *by = nullptr, *cy = nullptr;
  unsigned char* tmp = nullptr;
  int r = upload_table(tx, &bx, &cx);
  if (r == PF_OK) r = upload_table(ty, &by, &cy);
  if (r == PF_OK && cudaMalloc(&tmp, (size_t)H * new_w * 3) != cudaSuccess) r = fail(PF_ERR_CUDA, "pf_op_resize_u8: cudaMalloc");
  if (r == PF_OK) {
    // Pillow runs the horizontal pass first (over the rows the vertical pass needs: all of them here), each pass rounded to uint8
    cudaError_t le = (resize_u8_h_kernel<<<(unsigned)cdivl((long long)H * new_w, 256), 256, 0, st>>>(img, H, W, new_w, bx, cx, tx.ksize, tmp), cudaGetLastError());
    if (le == cudaSuccess) le = (resize_u8_v_kernel<<<(unsigned)cdivl((long long)new_h * new_w, 256), 256, 0, st>>>(tmp, H, new_w, new_h, by, cy, ty.ksize, out), cudaGetLastError());
    g_launches.fetch_add(2, std::memory_order_relaxed);
    if (le == cudaSuccess) le = cudaStreamSynchronize(st);
    if (le != cudaSuccess) r = fail(PF_ERR_CUDA, "pf_op_resize_u8: %s", cudaGetErrorString(le));
  }
  cudaFree(bx); cudaFree(cx); cudaFree(by); cudaFree(cy); cudaFree(tmp);
  return r;
}
int pf_op_resize_f32(const float* img, int H, int W, int C, int new_h, int new_w, float* out, void* stream) {
  if (!img || !out || H < 1 || W < 1 || C < 1 || new_h < 1 || new_w < 1) return fail(PF_ERR_ARG, "pf_op_resize_f32: bad argument");
  LAUNCHED((resize_f32_kernel<<<(unsigned)cdivl((long long)new_h * new_w * C, 256), 256, 0, (cudaStream_t)stream>>>(img, H, W, C, new_h, new_w, out), cudaGetLastError()));
  return PF_OK;
}
int pf_op_argmax_decode(const float* logits, float* field, int B, int HW, int NC, int is_gravity, void* stream) {
  if (!logits || !field || B < 1 || HW < 1 || NC < 1) return fail(PF_ERR_ARG, "pf_op_argmax_decode: bad argument");
  LAUNCHED((argmax_decode_kernel<<<(unsigned)cdivl((long long)B * HW, 256), 256, 0, (cudaStream_t)stream>>>(logits, field, B, HW, NC, is_gravity), cudaGetLastError()));
  return PF_OK;
}
int pf_op_pred_argmax_decode(const float* feat, int ld, int coff, const float* w, const float* bias, float* field, int B, int HW, int NC, int is_gravity,
                             void* stream) {
  if (!feat || !w || !bias || !field || B < 1 || HW < 1 || NC < 1 || NC > 256 || (ld & 3) || (coff & 3)) return fail(PF_ERR_ARG, "pf_op_pred_argmax_decode: bad argument");
  LAUNCHED((pred_argmax_decode_kernel<<<ew_grid((long long)B * HW * 4), 256, NC * 37 * 4, (cudaStream_t)stream>>>(feat, ld, coff, w, bias, field, B, HW, NC, is_gravity),
            cudaGetLastError()));
  return PF_OK;
}
int pf_op_postprocess(const float* vec, const float* lat, int n, const int32_t* height, const int32_t* width, float* gravity_original,
                      const int64_t* gravity_original_offset, float* latitude_original, const int64_t* latitude_original_offset, int lat_is_sin,
                      void* stream) {
  if (!vec || !lat || n < 1 || !height || !width || !gravity_original || !gravity_original_offset || !latitude_original || !latitude_original_offset)
    return fail(PF_ERR_ARG, "pf_op_postprocess: bad argument");
  PostImage* d_post = nullptr;
  CU(cudaMalloc(&d_post, n * sizeof(PostImage)));
  int r = launch_postprocess(vec, lat, n, height, width, gravity_original_offset, latitude_original_offset, gravity_original, latitude_original,
                             lat_is_sin, d_post, (cudaStream_t)stream);
  cudaError_t se = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(d_post);
  if (r == PF_OK && se != cudaSuccess) r = fail(PF_ERR_CUDA, "pf_op_postprocess: %s", cudaGetErrorString(se));
  return r;
}

int pf_op_fill_stream(float* dst, int64_t numel, float value, void* stream) {
  if (!dst || numel < 4 || (numel & 3) || ((uintptr_t)dst & 15)) return fail(PF_ERR_ARG, "pf_op_fill_stream: bad argument");
  LAUNCHED((fill_stream_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(dst), numel / 4, value), cudaGetLastError()));
  return PF_OK;
}
int pf_op_layernorm(const float* x, float* y, int64_t rows, int C, const float* w, const float* b, float eps, void* stream) {
  LAUNCHED(layernorm_launch(x, y, rows, C, w, b, eps, (cudaStream_t)stream));
  return PF_OK;
}
int pf_op_attention(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream) {
  TRY(configure_current_device());
  if (C != heads * kAttnD) return fail(PF_ERR_ARG, "pf_op_attention: head_dim must be 64");
  LAUNCHED(attention_launch(q, kv, out, B, N, C, heads, (cudaStream_t)stream));
  return PF_OK;
}
int pf_op_attention_mma(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream) {
  TRY(configure_current_device());
  if (C != heads * kAmD) return fail(PF_ERR_ARG, "pf_op_attention_mma: head_dim must be 64");
  LAUNCHED(attention_mma_launch(q, kv, out, B, N, C, heads, (cudaStream_t)stream));
  return PF_OK;
}
int pf_op_attention_tc(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream) {
  if (!q || !kv || !out || C != heads * kAtcD) return fail(PF_ERR_ARG, "pf_op_attention_tc: head_dim must be 64");
  TRY(configure_current_device());
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  pf_engine tmp;
  tmp.device = dev;
  tmp.sm_count = prop.multiProcessorCount;
  const long long nq = (long long)B * N * C, nkv = (long long)B * kAtcKeys * 2 * C;
  char* scratch = nullptr;
  CU(cudaMalloc(&scratch, (2 * nq + nkv) * 4 + 8192));
  Fwd F{&tmp, Arena{}, st, false, B};
  F.ar.base = scratch; F.ar.cap = (2 * nq + nkv) * 4 + 8192;
  SplitT qs = F.salloc((long long)B * N, C), kvs = F.salloc((long long)B * kAtcKeys, 2 * C), as = F.salloc((long long)B * N, C);
  int r = PF_OK;
  cudaError_t le = (split_kernel<<<(unsigned)cdivl(nq, 256), 256, 0, st>>>(q, qs.hi, qs.lo, nq, 0), cudaGetLastError());
  if (le == cudaSuccess) le = (split_kernel<<<(unsigned)cdivl(nkv, 256), 256, 0, st>>>(kv, kvs.hi, kvs.lo, nkv, 0), cudaGetLastError());
  if (le != cudaSuccess) r = fail(PF_ERR_CUDA, "split_kernel: %s", cudaGetErrorString(le));
  if (r == PF_OK) r = F.attention_tc(qs, kvs, as, B, N, C, heads);
  if (r == PF_OK) {
    le = (merge_split_kernel<<<(unsigned)cdivl(nq, 256), 256, 0, st>>>(as.hi, as.lo, out, nq), cudaGetLastError());
    if (le != cudaSuccess) r = fail(PF_ERR_CUDA, "merge_split_kernel: %s", cudaGetErrorString(le));
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(scratch);
  if (r == PF_OK && se != cudaSuccess) r = fail(PF_ERR_CUDA, "pf_op_attention_tc: %s", cudaGetErrorString(se));
  return r;
}
int pf_op_dwconv3x3_gelu(const float* x, float* y, int B, int H, int W, int C, const float* w, const float* bias, void* stream) {
  if (C % 4) return fail(PF_ERR_ARG, "C %% 4");
  LAUNCHED(launch_pdl(dwconv3x3_gelu_kernel, dim3(ew_grid((long long)B * ((H + 1) / 2) * ((W + PF_DW3_PX - 1) / PF_DW3_PX) * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, y, B, H, W, C, w, bias, nullptr, nullptr));
  return PF_OK;
}
int pf_op_dwconv7x7(const float* x, float* y, int B, int H, int W, int C, const float* w, const float* bias, void* stream) {
  if (C % 4) return fail(PF_ERR_ARG, "C %% 4");
  LAUNCHED(launch_pdl(dwconv7x7_kernel, dim3(ew_grid((long long)B * ((H + 1) / 2) * ((W + PF_DW7_PX - 1) / PF_DW7_PX) * (C / 4))), dim3(256), 0, (cudaStream_t)stream, x, y, B, H, W, C, w, bias));
  return PF_OK;
}
int pf_op_upsample2x(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  if (C % 4) return fail(PF_ERR_ARG, "C %% 4");
  LAUNCHED(launch_pdl(upsample2x_kernel, dim3(ew_grid(upsample2x_threads(B, H, W, C))), dim3(256), 0, (cudaStream_t)stream, x, C, 0, y, C, 0, B, H, W, C, nullptr, nullptr));
  return PF_OK;
}
int pf_op_preprocess(const uint8_t* img, int H, int W, const float* mean3, const float* std3, float* y, void* stream) {
  TRY(configure_current_device());
  // standalone tables (not cached): test entry point only
  ResampleTable tx = make_resample_table(W, kNet), ty = make_resample_table(H, kNet);
  if (ty.ksize + 1 > kPreMaxSmemRows) return fail(PF_ERR_ARG, "image too tall");
  int *bx, *cx, *by, *cy;
  PreImage* d;
  CU(cudaMalloc(&bx, tx.bounds.size() * 4)); CU(cudaMalloc(&cx, tx.coeffs.size() * 4));
  CU(cudaMalloc(&by, ty.bounds.size() * 4)); CU(cudaMalloc(&cy, ty.coeffs.size() * 4));
  CU(cudaMalloc(&d, sizeof(PreImage)));
  CU(cudaMemcpy(bx, tx.bounds.data(), tx.bounds.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(cx, tx.coeffs.data(), tx.coeffs.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(by, ty.bounds.data(), ty.bounds.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(cy, ty.coeffs.data(), ty.coeffs.size() * 4, cudaMemcpyHostToDevice));
  PreImage pi{0, H, W, tx.ksize, ty.ksize, bx, cx, by, cy};
  CU(cudaMemcpy(d, &pi, sizeof pi, cudaMemcpyHostToDevice));
  int rows = pre_rows_needed(H);
  if (rows > kPreMaxSmemRows) rows = kPreMaxSmemRows;
  const int smem = rows * kNet * 3;
  LAUNCHED((preprocess_kernel<<<dim3(kNet / kPreRows, 1), kNet, smem, (cudaStream_t)stream>>>(img, d, y, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], rows),
            cudaGetLastError()));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  cudaFree(bx); cudaFree(cx); cudaFree(by); cudaFree(cy); cudaFree(d);
  return PF_OK;
}

}  // extern "C"
