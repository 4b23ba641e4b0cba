// C ABI (include/str2str_b200.h): context, weight preparation, workspace, and the network forward that
// strings the kernels together.  Host-side only orchestration; every arithmetic op is a kernel of this library.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/str2str_b200.h"
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

long long g_launch_count = 0;
bool g_profile_on = false;
bool g_pdl = [] { const char* e = getenv("S2S_PDL"); return e ? atoi(e) != 0 : true; }();
static thread_local std::string g_error;

struct ProfRec { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; double ms = 0; long long n = 0; };
static std::map<std::string, ProfRec> g_prof;

ProfScope::ProfScope(const char* n, cudaStream_t s) : name(n), st(s) {
  if (!g_profile_on) return;
  cudaEventCreate(&e0);
  cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEvent_t e1;
  cudaEventCreate(&e1);
  cudaEventRecord(e1, st);
  g_prof[name].ev.emplace_back(e0, e1);
}
const char* prof_intern(const std::string& name) {
  static std::map<std::string, int> names;
  return names.emplace(name, 0).first->first.c_str();
}
static void prof_drain() {
  for (auto& kv : g_prof) {
    for (auto& pr : kv.second.ev) {
      cudaEventSynchronize(pr.second);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pr.first, pr.second);
      kv.second.ms += ms;
      kv.second.n += 1;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    kv.second.ev.clear();
  }
}

namespace {

// ---- weight preparation kernels ----------------------------------------------------------------------
// dst[c][r] = bf16(src[r][c0 + c]) : transposed bf16 image of a weight sub-block
__global__ void prep_bf16_t_kernel(const float* __restrict__ src, long ld, int rows, int c0, int cols, bf16* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(long)c * rows + r] = __float2bfloat16_rn(src[(long)r * ld + c0 + c]);
}
// hi[r][c] = bf16(src[r][c0+c]); lo = bf16(src - hi)   (lo may be null)
__global__ void prep_bf16_split_kernel(const float* __restrict__ src, long ld, int rows, int c0, int cols,
                                       bf16* __restrict__ hi, bf16* __restrict__ lo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float v = src[(long)r * ld + c0 + c];
  const bf16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// dst[c][r] = src[r][c0 + c]  (fp32 transpose of a sub-block)
__global__ void prep_f32_t_kernel(const float* __restrict__ src, long ld, int rows, int c0, int cols, float* __restrict__ dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(long)c * rows + r] = src[(long)r * ld + c0 + c];
}

struct Slab {  // bump allocator over one cudaMalloc
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = 0;
  }
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    S2S_CHECK(off <= cap, "workspace overflow");
    return p;
  }
};

struct IpaW {
  float *proj_w, *proj_b;  // [6816][256], [6816]: q | kv | q_points | kv_points
  bf16 *Wb_hi, *Wb_lo, *wb_img;
  float *Wdz_t, *pt_w;
};
struct EtW {
  bf16 *W1z, *W2, *Wfh, *Wfz, *W1zt, *W2t, *Wft, *Wfzt;
  bf16* wimg3;  // tcgen05 weight image (pre-swizzled blocks in consumption order), wimg_copies replicas
};

}  // namespace
}  // namespace s2s

using namespace s2s;

struct s2s_ctx {
  std::map<std::string, std::pair<const float*, int64_t>> params;
  bool finalized = false;
  int opt_pair = 1, opt_node = 1, opt_ipa = 1, opt_table = 1, opt_et_pair = 1;
  int opt_fork = 1;   // few residue rows: independent kernels of a block run on a side stream (Fork below); S2S_FORK=0 disables
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork[4] = {}, ev_join[4] = {};
  int opt_chain = 1;  // row-local layers of the node track fused into gemm_chain launches: 0 never, 1 when the row panels outnumber half the SMs, 2 always (do_trunk)
  int tfm_passes = 1;  // sequence-transformer in_proj + attention GEMMs: 1 = single bf16 (default; trajectory error unchanged, tools/traj_parity.py), 3 = split-bf16 (S2S_TFM_PASSES=3)
  int wimg_copies = 8;  // replicated EdgeTransition weight images (set S2S_WIMG_COPIES to override)
  int cur_prec = 0;  // default precision class of linear() calls (0 exact, 3 split-bf16, 1 bf16); set per stage
  float *tfreq = nullptr, *pdenom = nullptr, *bin_lower = nullptr, *backbone = nullptr;
  Slab wslab;  // derived weights
  IpaW ipa[N_BLK];
  EtW et[N_BLK - 1];
  bf16 *ee_W2, *ee_W3, *ee_W2t, *ee_W3t, *ee_wimg;
  float* ee_Wd;
  float* ee_vec;  // [b2 | b3 | ln_w | ln_b] packed (4 x 128): one copy into the pipelined embedder's constant bank per call
  // workspace
  Slab ws;
  int cap_B = 0, cap_L = 0, d_min = 0, n_off = 0, cap_off = 0;  // cap_L is a padded length (pad_len)
  // internal chain-length padding: padded copies [B][Lp] of the per-residue inputs / outputs; `hard` marks real rows
  float *pd_rig, *pd_sc, *pd_rmask, *pd_fixed, *pd_psi, *pd_hard, *pd_orig, *pd_opsi;
  long long* pd_ridx;
  const float* hard = nullptr;  // non-null while a padded call is running
  // edge-embedder table (pair_tc4.cu MODE 2): allocated when it can pay for the reserved shape
  bf16* ee_table = nullptr; size_t ee_table_elems = 0;
  int* ee_ctl = nullptr;
  unsigned char* ee_cls = nullptr;
  float *feat65, *tf33, *node, *init_node, *a256, *b256, *proj, *feats, *q_pts, *k_pts, *v_pts, *S, *opt;
  float *skip_w_all = nullptr, *skip_b_all = nullptr;  // the four skip_embed layers stacked: [4 * 64][256], [4 * 64]
  float* skip_all = nullptr;                            // [R][4 * 64]: skip_embed_b(init_node) of every block, computed once
  float *skip64, *x320, *t320, *y320, *qkv, *nprime, *u384, *v384, *p128, *q128, *Ti, *Tj, *Tpos, *relfeat;
  float *quat, *trans, *upd6, *psi_u, *diffuse, *keybias;
  bf16 *z, *nprime_bf16;
  // tensor-core node track: bf16 hi/lo images of every weight matrix, scratch split buffers, attention operands
  std::map<const float*, std::pair<size_t, std::pair<bf16*, bf16*>>> wsplit;  // fp32 base -> (numel, (hi, lo))
  bf16 *sa_hi, *sa_lo, *qkv_bf16, *P_bf16;
  bf16 *qp_aug, *kp_aug, *vp_hi, *vp_lo, *P_lo;  // fused-logits / split-P operands of the second-generation IPA path
  float* colbias;
  bf16 *tq_hi, *tq_lo, *tP_lo;  // sequence-transformer attention operands (split bf16; v is read in place, MN-major)
  // split-bf16 companions of the node-track activations, written by the producing kernel's epilogue
  bf16 *node_hi, *node_lo, *init_hi, *init_lo, *a256_hi, *a256_lo, *b256_hi, *b256_lo;
  bf16 *x320_hi, *x320_lo, *t320_hi, *t320_lo, *y320_hi, *y320_lo, *nprime_hi, *nprime_lo;

  const float* P(const std::string& n) const {
    auto it = params.find(n);
    S2S_CHECK(it != params.end(), "missing parameter " + n);
    return it->second.first;
  }
};

namespace {

enum Prec { EXACT = 0, TC3 = 3, TC1 = 1 };
// RAII marker of a padded call: c->hard is non-null exactly while the padded stages run (also on error paths)
struct HardScope {
  s2s_ctx* c;
  HardScope(s2s_ctx* ctx, const float* hard) : c(ctx) { c->hard = hard; }
  ~HardScope() { c->hard = nullptr; }
};

struct PrecScope {
  s2s_ctx* c;
  PrecScope(s2s_ctx* ctx, int p) : c(ctx) { c->cur_prec = p; }
  ~PrecScope() { c->cur_prec = EXACT; }
};

struct Split { bf16* hi = nullptr; bf16* lo = nullptr; };  // split-bf16 image of an fp32 activation (dense, pitch = width)

// A few launches on the context's side stream, concurrent with what follows on `st` until join().  With few residue rows a kernel
// occupies a fraction of the SMs and the forward is a chain of launch latencies (~10 us per dependent tensor-core GEMM), so kernels
// that do not depend on each other — the q|k|v projection and the point projection, P.v and P.v_pts, the per-residue
// EdgeTransition terms, the frame update — overlap instead.  Event record / wait pairs: legal under CUDA-graph capture, where
// they become parallel branches of the graph.  Off (everything on `st`) when the kernels fill the GPU anyway.
struct Fork {
  s2s_ctx* c;
  cudaStream_t st;
  int slot;
  bool on;
  Fork(s2s_ctx* ctx, cudaStream_t main, int slot_, bool enable) : c(ctx), st(main), slot(slot_), on(enable && ctx->opt_fork && ctx->side && !g_profile_on) {
    if (on) {
      S2S_CUDA(cudaEventRecord(c->ev_fork[slot], st));
      S2S_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork[slot], 0));
    }
  }
  cudaStream_t stream() const { return on ? c->side : st; }
  void join() {
    if (!on) return;
    on = false;
    S2S_CUDA(cudaEventRecord(c->ev_join[slot], c->side));
    S2S_CUDA(cudaStreamWaitEvent(st, c->ev_join[slot], 0));
  }
  ~Fork() {
    if (on) {  // error path: never leave the side stream detached from a capture
      cudaEventRecord(c->ev_join[slot], c->side);
      cudaStreamWaitEvent(st, c->ev_join[slot], 0);
    }
  }
};
bool few_rows(int R) { return 2 * ceil_div(R, 128) <= sm_count(); }

// bf16 (hi, lo) image of a weight (or of a sub-block of one) registered at finalize
std::pair<const bf16*, const bf16*> weight_split(const s2s_ctx* c, const float* W) {
  auto it = c->wsplit.upper_bound(W);
  S2S_CHECK(it != c->wsplit.begin(), "weight_split: unknown weight pointer");
  --it;
  const size_t off = (size_t)(W - it->first);
  S2S_CHECK(off < it->second.first, "weight_split: pointer outside any registered weight");
  return {it->second.second.first + off, it->second.second.second + off};
}

// y = act((x W^T * row_pre + bias)) * row_post + res.  prec selects exact fp32 FFMA or the tensor-core GEMM
// (3-pass split-bf16 / single bf16) when the context runs with node_gemm = 1.
void linear(const s2s_ctx* c, const float* x, long ldx, const float* W, long ldw, const float* bias, float* y, long ldy,
            int M, int N, int K, cudaStream_t st, int relu = 0, const float* res = nullptr, long ldres = 0,
            const float* row_pre = nullptr, const float* row_post = nullptr, int prec_arg = -1, Split in = Split(),
            Split out = Split()) {
  const Prec prec = (Prec)(prec_arg < 0 ? c->cur_prec : prec_arg);
  if (prec != EXACT && c->opt_node == 1 && K % 16 == 0 && ldx % 4 == 0 && ldw % 8 == 0) {
    if (!in.hi) {  // no image from the producer: split here
      split_bf16(x, ldx, M, K, c->sa_hi, prec == TC3 ? c->sa_lo : nullptr, st);
      in.hi = c->sa_hi; in.lo = c->sa_lo;
    }
    const auto w = weight_split(c, W);
    TcGemm g;
    g.A_hi = in.hi; g.A_lo = in.lo; g.a_rows = M; g.a_cols = K; g.a_pitch = K;
    g.out_hi = out.hi; g.out_lo = out.lo; g.ldo = N;
    g.B_hi = w.first; g.B_lo = w.second; g.b_rows = N; g.b_cols = K; g.b_pitch = ldw;
    g.M = M; g.N = N; g.K = K; g.passes = (int)prec; g.relu = relu;
    g.bias = bias; g.row_pre = row_pre; g.row_post = row_post; g.res = res; g.ldres = ldres;
    g.C = y; g.ldc = ldy;
    gemm_tc(g, st);
    return;
  }
  GemmArgs g;
  g.A = x; g.lda = ldx; g.B = W; g.ldb = ldw; g.C = y; g.ldc = ldy;
  g.bias = bias; g.res = res; g.ldres = ldres; g.row_pre = row_pre; g.row_post = row_post;
  g.M = M; g.N = N; g.K = K; g.relu = relu;
  gemm_f32(g, st);
  if (out.hi) split_bf16(y, ldy, M, N, out.hi, out.lo, st);  // exact path taken although an image was requested
}

struct ParamSpec { std::string name; int64_t numel; };

std::vector<ParamSpec> expected_params() {
  std::vector<ParamSpec> v;
  auto lin = [&](const std::string& n, int o, int i) { v.push_back({n + ".weight", (int64_t)o * i}); v.push_back({n + ".bias", o}); };
  auto ln = [&](const std::string& n, int d) { v.push_back({n + ".weight", d}); v.push_back({n + ".bias", d}); };
  lin("embedder.node_embed.0", 256, 65); lin("embedder.node_embed.2", 256, 256); lin("embedder.node_embed.4", 256, 256); ln("embedder.node_embed.5", 256);
  lin("embedder.edge_embed.0", 128, 120); lin("embedder.edge_embed.2", 128, 128); lin("embedder.edge_embed.4", 128, 128); ln("embedder.edge_embed.5", 128);
  for (int b = 0; b < N_BLK; ++b) {
    const std::string t = "translator.trunk.", s = std::to_string(b);
    v.push_back({t + "ipa_" + s + ".head_weights", 8});
    lin(t + "ipa_" + s + ".linear_q", 2048, 256); lin(t + "ipa_" + s + ".linear_kv", 4096, 256);
    lin(t + "ipa_" + s + ".linear_q_points", 192, 256); lin(t + "ipa_" + s + ".linear_kv_points", 480, 256);
    lin(t + "ipa_" + s + ".linear_b", 8, 128); lin(t + "ipa_" + s + ".down_z", 32, 128);
    lin(t + "ipa_" + s + ".linear_out", 256, IPA_FEAT);
    ln(t + "ipa_ln_" + s, 256); lin(t + "skip_embed_" + s, 64, 256);
    for (int l = 0; l < 2; ++l) {
      const std::string tl = t + "transformer_" + s + ".layers." + std::to_string(l) + ".";
      v.push_back({tl + "self_attn.in_proj_weight", 960 * 320}); v.push_back({tl + "self_attn.in_proj_bias", 960});
      lin(tl + "self_attn.out_proj", 320, 320); lin(tl + "linear1", 320, 320); lin(tl + "linear2", 320, 320);
      ln(tl + "norm1", 320); ln(tl + "norm2", 320);
    }
    lin(t + "linear_" + s, 256, 320);
    lin(t + "node_transition_" + s + ".linear_1", 256, 256); lin(t + "node_transition_" + s + ".linear_2", 256, 256);
    lin(t + "node_transition_" + s + ".linear_3", 256, 256); ln(t + "node_transition_" + s + ".ln", 256);
    lin(t + "bb_update_" + s + ".linear", 6, 256);
    if (b < N_BLK - 1) {
      const std::string e = t + "edge_transition_" + s + ".";
      lin(e + "initial_embed", 128, 256); lin(e + "trunk.0", 384, 384); lin(e + "trunk.2", 384, 384);
      lin(e + "final_layer", 128, 384); ln(e + "layer_norm", 128);
    }
  }
  lin("translator.torsion_pred.linear_1", 256, 256); lin("translator.torsion_pred.linear_2", 256, 256);
  lin("translator.torsion_pred.linear_3", 256, 256); lin("translator.torsion_pred.linear_final", 2, 256);
  return v;
}

void prep_t_bf16(const float* src, long ld, int rows, int c0, int cols, bf16* dst, cudaStream_t st) {
  prep_bf16_t_kernel<<<ceil_div((long)rows * cols, 256), 256, 0, st>>>(src, ld, rows, c0, cols, dst);
  S2S_LAUNCH_CHECK();
}
void prep_split(const float* src, long ld, int rows, int c0, int cols, bf16* hi, bf16* lo, cudaStream_t st) {
  prep_bf16_split_kernel<<<ceil_div((long)rows * cols, 256), 256, 0, st>>>(src, ld, rows, c0, cols, hi, lo);
  S2S_LAUNCH_CHECK();
}
void prep_t_f32(const float* src, long ld, int rows, int c0, int cols, float* dst, cudaStream_t st) {
  prep_f32_t_kernel<<<ceil_div((long)rows * cols, 256), 256, 0, st>>>(src, ld, rows, c0, cols, dst);
  S2S_LAUNCH_CHECK();
}

void do_finalize(s2s_ctx* c, cudaStream_t st) {
  for (const auto& ps : expected_params()) {
    auto it = c->params.find(ps.name);
    S2S_CHECK(it != c->params.end(), "missing parameter " + ps.name);
    S2S_CHECK(it->second.second == ps.numel, "parameter " + ps.name + " has " + std::to_string(it->second.second) +
                                                 " elements, expected " + std::to_string(ps.numel));
  }
  c->wslab.release();
  c->wslab.cap = 256u << 20;
  S2S_CUDA(cudaMalloc(&c->wslab.base, c->wslab.cap));
  const std::string t = "translator.trunk.";
  for (int b = 0; b < N_BLK; ++b) {
    const std::string ip = t + "ipa_" + std::to_string(b) + ".";
    IpaW& w = c->ipa[b];
    w.proj_w = c->wslab.take<float>(6816 * 256);
    w.proj_b = c->wslab.take<float>(6816);
    const struct { const char* n; int rows; int off; } parts[4] = {
        {"linear_q", 2048, 0}, {"linear_kv", 4096, 2048}, {"linear_q_points", 192, 6144}, {"linear_kv_points", 480, 6336}};
    for (const auto& p : parts) {
      S2S_CUDA(cudaMemcpyAsync(w.proj_w + (size_t)p.off * 256, c->P(ip + p.n + ".weight"), (size_t)p.rows * 256 * 4, cudaMemcpyDeviceToDevice, st));
      S2S_CUDA(cudaMemcpyAsync(w.proj_b + p.off, c->P(ip + p.n + ".bias"), (size_t)p.rows * 4, cudaMemcpyDeviceToDevice, st));
    }
    w.Wb_hi = c->wslab.take<bf16>(8 * 128);
    w.Wb_lo = c->wslab.take<bf16>(8 * 128);
    prep_split(c->P(ip + "linear_b.weight"), 128, 8, 0, 128, w.Wb_hi, w.Wb_lo, st);
    w.wb_img = c->wslab.take<bf16>(ipa_wb_img_elems());
    build_ipa_wb_img(c->P(ip + "linear_b.weight"), w.wb_img, st);
    w.Wdz_t = c->wslab.take<float>(128 * 32);
    prep_t_f32(c->P(ip + "down_z.weight"), 128, 32, 0, 128, w.Wdz_t, st);
    w.pt_w = c->wslab.take<float>(8);
    softplus_point_weights(c->P(ip + "head_weights"), w.pt_w, st);
    if (b < N_BLK - 1) {
      const std::string e = t + "edge_transition_" + std::to_string(b) + ".";
      EtW& x = c->et[b];
      const float *W1 = c->P(e + "trunk.0.weight"), *W2 = c->P(e + "trunk.2.weight"), *Wf = c->P(e + "final_layer.weight");
      x.W1z = c->wslab.take<bf16>(384 * 128); x.W1zt = c->wslab.take<bf16>(384 * 128);
      x.W2 = c->wslab.take<bf16>(384 * 384);  x.W2t = c->wslab.take<bf16>(384 * 384);
      x.Wfh = c->wslab.take<bf16>(128 * 384); x.Wft = c->wslab.take<bf16>(128 * 384);
      x.Wfz = c->wslab.take<bf16>(128 * 128); x.Wfzt = c->wslab.take<bf16>(128 * 128);
      prep_split(W1, 384, 384, 0, 128, x.W1z, nullptr, st);  prep_t_bf16(W1, 384, 384, 0, 128, x.W1zt, st);
      prep_split(W2, 384, 384, 0, 384, x.W2, nullptr, st);   prep_t_bf16(W2, 384, 384, 0, 384, x.W2t, st);
      prep_split(Wf, 384, 128, 0, 384, x.Wfh, nullptr, st);  prep_t_bf16(Wf, 384, 128, 0, 384, x.Wft, st);
      prep_split(Wf, 384, 128, 0, 128, x.Wfz, nullptr, st);  prep_t_bf16(Wf, 384, 128, 0, 128, x.Wfzt, st);
      x.wimg3 = c->wslab.take<bf16>(et3_wimg_elems() * c->wimg_copies);
      build_et3_wimg(W1, W2, Wf, x.wimg3, st);
      for (int k = 1; k < c->wimg_copies; ++k)
        S2S_CUDA(cudaMemcpyAsync(x.wimg3 + k * et3_wimg_elems(), x.wimg3, et3_wimg_elems() * 2, cudaMemcpyDeviceToDevice, st));
    }
  }
  {
    const float *W1 = c->P("embedder.edge_embed.0.weight"), *W2 = c->P("embedder.edge_embed.2.weight"),
                *W3 = c->P("embedder.edge_embed.4.weight");
    c->ee_W2 = c->wslab.take<bf16>(128 * 128); c->ee_W2t = c->wslab.take<bf16>(128 * 128);
    c->ee_W3 = c->wslab.take<bf16>(128 * 128); c->ee_W3t = c->wslab.take<bf16>(128 * 128);
    c->ee_Wd = c->wslab.take<float>(N_BINS * 128);
    prep_split(W2, 128, 128, 0, 128, c->ee_W2, nullptr, st); prep_t_bf16(W2, 128, 128, 0, 128, c->ee_W2t, st);
    prep_split(W3, 128, 128, 0, 128, c->ee_W3, nullptr, st); prep_t_bf16(W3, 128, 128, 0, 128, c->ee_W3t, st);
    prep_t_f32(W1, 120, 128, 98, N_BINS, c->ee_Wd, st);
    c->ee_wimg = c->wslab.take<bf16>(ee_wimg_elems());
    build_ee_wimg(W2, W3, c->ee_wimg, st);
    c->ee_vec = c->wslab.take<float>(4 * 128);
    const char* vn[4] = {"embedder.edge_embed.2.bias", "embedder.edge_embed.4.bias", "embedder.edge_embed.5.weight", "embedder.edge_embed.5.bias"};
    for (int v = 0; v < 4; ++v) S2S_CUDA(cudaMemcpyAsync(c->ee_vec + v * 128, c->P(vn[v]), 128 * 4, cudaMemcpyDeviceToDevice, st));
  }
  // the skip connections of all four blocks read the same init_node (ipa.py:353-356): one stacked GEMM serves them
  c->skip_w_all = c->wslab.take<float>(N_BLK * D_SKIP * 256);
  c->skip_b_all = c->wslab.take<float>(N_BLK * D_SKIP);
  for (int b = 0; b < N_BLK; ++b) {
    const std::string sk = t + "skip_embed_" + std::to_string(b);
    S2S_CUDA(cudaMemcpyAsync(c->skip_w_all + (size_t)b * D_SKIP * 256, c->P(sk + ".weight"), (size_t)D_SKIP * 256 * 4, cudaMemcpyDeviceToDevice, st));
    S2S_CUDA(cudaMemcpyAsync(c->skip_b_all + b * D_SKIP, c->P(sk + ".bias"), D_SKIP * 4, cudaMemcpyDeviceToDevice, st));
  }
  // bf16 hi/lo images of every matrix the tensor-core node track multiplies by
  c->wsplit.clear();
  auto reg = [&](const float* W, size_t numel, int cols) {
    bf16* hi = c->wslab.take<bf16>(numel);
    bf16* lo = c->wslab.take<bf16>(numel);
    prep_split(W, cols, (int)(numel / cols), 0, cols, hi, lo, st);
    c->wsplit[W] = {numel, {hi, lo}};
  };
  for (const auto& ps : expected_params()) {
    const std::string& n = ps.name;
    if (n.size() < 7 || n.compare(n.size() - 6, 6, "weight") != 0 || ps.numel < 4096) continue;
    if (n.find("in_proj") != std::string::npos) { reg(c->P(n), ps.numel, 320); continue; }
    if (n.find("ipa_") != std::string::npos && (n.find("linear_q") != std::string::npos || n.find("linear_kv") != std::string::npos)) continue;
    int cols = 256;
    if (n.find("transformer") != std::string::npos || (n.find("trunk.linear_") != std::string::npos)) cols = 320;
    if (n.find("linear_out") != std::string::npos) cols = IPA_FEAT;
    if (n.find("trunk.0") != std::string::npos || n.find("trunk.2") != std::string::npos || n.find("final_layer") != std::string::npos) cols = 384;
    const bool ne_tc = n.find("node_embed.2.") != std::string::npos || n.find("node_embed.4.") != std::string::npos;  // optional TC3 path
    if ((n.find("embedder") != std::string::npos && !ne_tc) || n.find("linear_b.") != std::string::npos || n.find("down_z") != std::string::npos) continue;
    reg(c->P(n), ps.numel, cols);
  }
  for (int b = 0; b < N_BLK; ++b) reg(c->ipa[b].proj_w, (size_t)6816 * 256, 256);
  reg(c->skip_w_all, (size_t)N_BLK * D_SKIP * 256, 256);
  // Wfh above holds the full [128][384] final-layer image; only its action on h2 (all 384 inputs) is used:
  // final_layer(h2 + x) = Wf h2 + Wf[:, :128] z + Wf[:,128:256] n_i + Wf[:,256:] n_j.
  c->finalized = true;
}

// Chain lengths run on the tensor-core tiling: the pair kernels tile the flattened pair tensor in 32-row segments that must
// not straddle an i row, the batched attention GEMMs need K = L % 16 == 0.  Other lengths are padded INSIDE the library
// (pad_inputs / repitch_* in rows.cu) and come back un-padded; results equal the un-padded computation because appended
// residues are excluded from every reduction over residues (see rows.cu: masks_kernel, pad_inputs_kernel).
int pad_len(int L) { return (L + 31) & ~31; }

void do_reserve(s2s_ctx* c, int B, int L_user, int d_min, int d_max, cudaStream_t st) {
  S2S_CHECK(c->finalized, "s2s_finalize must run before s2s_reserve");
  S2S_CHECK(B > 0 && L_user > 0, "reserve: bad shape");
  const int L = pad_len(L_user);
  const bool keep_table = d_max < d_min;  // shape only: the relative-position table stays as planned (or the 1-row default)
  if (keep_table) {
    d_min = c->ws.base ? c->d_min : 0;
    d_max = c->ws.base ? c->d_min + c->n_off - 1 : 0;
  }
  const int n_off = d_max - d_min + 1;
  if (!(B <= c->cap_B && L <= c->cap_L && n_off <= c->cap_off && c->ws.base)) {
    if (c->ws.base) S2S_CUDA(cudaDeviceSynchronize());
    c->ws.release();
    B = std::max(B, c->cap_B);  // never shrink: an engine alternating between shapes settles on one allocation
    const int Lc = std::max(L, c->cap_L), cap_off = std::max(n_off, c->cap_off);
    const size_t R = (size_t)B * Lc;
    size_t bytes = 0;
    auto add = [&](size_t n, size_t sz) { bytes += ((n * sz + 255) & ~size_t(255)) + 256; };
    add(R * 65, 4); add(R * 33, 4); for (int k = 0; k < 4; ++k) add(R * 256, 4);
    add(R * 6816, 4); add(R * IPA_FEAT, 4); add(R * 192, 4); add(R * 192, 4); add(R * 288, 4);
    add((size_t)B * N_H * Lc * Lc, 4); add(R * 288, 4);
    add(R * 64, 4); for (int k = 0; k < 3; ++k) add(R * 320, 4); add(R * 960, 4);
    add(R * N_BLK * D_SKIP, 4);
    add(R * 128, 4); add(R * 384, 4); add(R * 384, 4); add(R * 128, 4); add(R * 128, 4); add(R * 128, 4); add(R * 128, 4);
    add((size_t)cap_off * 128, 4); add((size_t)cap_off * 32, 4);
    add(R * 4, 4); add(R * 3, 4); add(R * 6, 4); add(R * 2, 4); add(R, 4); add(R, 4);
    add(R * Lc * C_Z, 2); add(R * 128, 2);
    add(R * IPA_FEAT, 2); add(R * IPA_FEAT, 2); add(R * 6144, 2); add(R * 2048, 2); add((size_t)B * N_H * Lc * Lc, 2);
    add(R * 960, 2); add(R * 960, 2); add(R * 320, 2); add(R * 320, 2); add((size_t)B * TFM_H * Lc * Lc, 2);
    for (int k = 0; k < 8; ++k) add(R * 256, 2);
    for (int k = 0; k < 6; ++k) add(R * 320, 2);
    add(R * 128, 2);
    add(R * N_H * PT_K, 2); add(R * N_H * PT_K, 2); add(R * N_H * VP_PITCH, 2); add(R * N_H * VP_PITCH, 2);
    add((size_t)B * N_H * Lc * Lc, 2); add(R * N_H, 4);
    add(R * 7, 4); add(R * 3, 4); add(R, 4); add(R, 4); add(R * 2, 4); add(R, 4); add(R * 7, 4); add(R * 2, 4); add(R, 8);
    // the embedding table pays when its rows (offsets x 23 slots) are well below the L^2 pair rows of a decoy
    const size_t tab_elems = ((size_t)cap_off * (N_BINS + 1) * 2 < (size_t)Lc * Lc) ? edge_embed_table_elems(B, cap_off, 4) : 0;
    add(tab_elems, 2); add(edge_embed_ctl_ints(B), 4); add(R, 1);
    c->ws.cap = bytes + 4096;
    S2S_CUDA(cudaMalloc(&c->ws.base, c->ws.cap));
    Slab& w = c->ws;
    c->feat65 = w.take<float>(R * 65); c->tf33 = w.take<float>(R * 33);
    c->node = w.take<float>(R * 256); c->init_node = w.take<float>(R * 256); c->a256 = w.take<float>(R * 256); c->b256 = w.take<float>(R * 256);
    c->proj = w.take<float>(R * 6816); c->feats = w.take<float>(R * IPA_FEAT);
    c->q_pts = w.take<float>(R * 192); c->k_pts = w.take<float>(R * 192); c->v_pts = w.take<float>(R * 288);
    c->S = w.take<float>((size_t)B * N_H * Lc * Lc); c->opt = w.take<float>(R * 288);
    c->skip_all = w.take<float>(R * N_BLK * D_SKIP);
    c->skip64 = w.take<float>(R * 64); c->x320 = w.take<float>(R * 320); c->t320 = w.take<float>(R * 320); c->y320 = w.take<float>(R * 320);
    c->qkv = w.take<float>(R * 960);
    c->nprime = w.take<float>(R * 128); c->u384 = w.take<float>(R * 384); c->v384 = w.take<float>(R * 384);
    c->p128 = w.take<float>(R * 128); c->q128 = w.take<float>(R * 128); c->Ti = w.take<float>(R * 128); c->Tj = w.take<float>(R * 128);
    c->Tpos = w.take<float>((size_t)cap_off * 128); c->relfeat = w.take<float>((size_t)cap_off * 32);
    c->quat = w.take<float>(R * 4); c->trans = w.take<float>(R * 3); c->upd6 = w.take<float>(R * 6); c->psi_u = w.take<float>(R * 2);
    c->diffuse = w.take<float>(R); c->keybias = w.take<float>(R);
    c->z = w.take<bf16>(R * Lc * C_Z);
    c->nprime_bf16 = w.take<bf16>(R * 128);
    c->sa_hi = w.take<bf16>(R * IPA_FEAT); c->sa_lo = w.take<bf16>(R * IPA_FEAT);
    c->qkv_bf16 = w.take<bf16>(R * 6144); c->P_bf16 = w.take<bf16>((size_t)B * N_H * Lc * Lc);
    c->tq_hi = w.take<bf16>(R * 960); c->tq_lo = w.take<bf16>(R * 960);
    c->tP_lo = w.take<bf16>((size_t)B * TFM_H * Lc * Lc);
    c->node_hi = w.take<bf16>(R * 256); c->node_lo = w.take<bf16>(R * 256); c->init_hi = w.take<bf16>(R * 256); c->init_lo = w.take<bf16>(R * 256);
    c->a256_hi = w.take<bf16>(R * 256); c->a256_lo = w.take<bf16>(R * 256); c->b256_hi = w.take<bf16>(R * 256); c->b256_lo = w.take<bf16>(R * 256);
    c->x320_hi = w.take<bf16>(R * 320); c->x320_lo = w.take<bf16>(R * 320); c->t320_hi = w.take<bf16>(R * 320); c->t320_lo = w.take<bf16>(R * 320);
    c->y320_hi = w.take<bf16>(R * 320); c->y320_lo = w.take<bf16>(R * 320);
    c->nprime_hi = c->nprime_bf16; c->nprime_lo = w.take<bf16>(R * 128);
    c->qp_aug = w.take<bf16>(R * N_H * PT_K); c->kp_aug = w.take<bf16>(R * N_H * PT_K);
    c->vp_hi = w.take<bf16>(R * N_H * VP_PITCH); c->vp_lo = w.take<bf16>(R * N_H * VP_PITCH);
    c->P_lo = w.take<bf16>((size_t)B * N_H * Lc * Lc); c->colbias = w.take<float>(R * N_H);
    c->pd_rig = w.take<float>(R * 7); c->pd_sc = w.take<float>(R * 3); c->pd_rmask = w.take<float>(R); c->pd_fixed = w.take<float>(R);
    c->pd_psi = w.take<float>(R * 2); c->pd_hard = w.take<float>(R); c->pd_orig = w.take<float>(R * 7); c->pd_opsi = w.take<float>(R * 2);
    c->pd_ridx = w.take<long long>(R);
    c->ee_table = tab_elems ? w.take<bf16>(tab_elems) : nullptr; c->ee_table_elems = tab_elems;
    c->ee_ctl = w.take<int>(edge_embed_ctl_ints(B)); c->ee_cls = w.take<unsigned char>(R);
    S2S_CUDA(cudaMemsetAsync(c->ee_ctl, 0, edge_embed_ctl_ints(B) * sizeof(int), st));  // the setup kernel's arrival counter starts at 0
    c->cap_B = B; c->cap_L = Lc; c->cap_off = cap_off;
    c->n_off = 0;  // the table memory is new: rebuild it below
  }
  if (keep_table && c->n_off == n_off && c->d_min == d_min) return;
  // relative-position table: Tpos[r] = W1[:,66:98] pos(d_min + r)   (denoising_ipa.py:144-149)
  c->d_min = d_min; c->n_off = n_off;
  relpos_features(c->pdenom, c->relfeat, d_min, n_off, st);
  linear(c, c->relfeat, 32, c->P("embedder.edge_embed.0.weight") + 66, 120, nullptr, c->Tpos, 128, n_off, 128, 32, st);
}

void check_shape(const s2s_ctx* c, int B, int L) {
  S2S_CHECK(c->finalized, "context not finalized");
  S2S_CHECK(c->ws.base && B <= c->cap_B && pad_len(L) <= c->cap_L && B > 0 && L > 0,
            "workspace too small: call s2s_reserve(B, L, ...) first");
}

// EmbeddingModule.forward + mask multiply
void do_embed(s2s_ctx* c, int B, int L, const float* t, const long long* ridx, const float* fixed, const float* sc_ca,
              const float* rmask, float* node_out, bf16* z_out, cudaStream_t st) {
  const int R = B * L;
  node_features(t, ridx, fixed, c->tfreq, c->pdenom, c->feat65, c->tf33, B, L, st);
  const std::string ne = "embedder.node_embed.", ee = "embedder.edge_embed.";
  linear(c, c->feat65, 65, c->P(ne + "0.weight"), 65, c->P(ne + "0.bias"), c->a256, 256, R, 256, 65, st, 1);
  // layers 2 and 3 (K = 256): split-bf16 tensor-core GEMMs (~16 mantissa bits per operand; measured on B200: trajectory error
  // 1.259e-5 vs 1.245e-5 with exact fp32 at L = 64 x 100 steps, 8.65e-6 vs 8.27e-6 at L = 128 x 50 — profiles/r01c_traj_parity.log).
  // Single bf16 is NOT enough here (6.7e-4, DESIGN.md precision table).  S2S_NODE_EMBED_TC=0 restores the exact FFMA GEMMs.
  static const int ne_env = [] { const char* e = getenv("S2S_NODE_EMBED_TC"); return e ? atoi(e) : 1; }();
  const int ne_prec = (ne_env && c->opt_node == 1) ? (int)TC3 : (int)EXACT;
  linear(c, c->a256, 256, c->P(ne + "2.weight"), 256, c->P(ne + "2.bias"), c->b256, 256, R, 256, 256, st, 1, nullptr, 0, nullptr, nullptr, ne_prec);
  linear(c, c->b256, 256, c->P(ne + "4.weight"), 256, c->P(ne + "4.bias"), c->a256, 256, R, 256, 256, st, 0, nullptr, 0, nullptr, nullptr, ne_prec);
  layernorm(c->a256, nullptr, c->P(ne + "5.weight"), c->P(ne + "5.bias"), rmask, node_out, R, 256, st);
  const float* W1 = c->P(ee + "0.weight");
  linear(c, c->tf33, 33, W1, 120, c->P(ee + "0.bias"), c->Ti, 128, R, 128, 33, st);
  linear(c, c->tf33, 33, W1 + 33, 120, nullptr, c->Tj, 128, R, 128, 33, st);
  EdgeEmbedArgs a;
  a.B = B; a.L = L; a.d_min = c->d_min; a.n_off = c->n_off;
  a.Ti = c->Ti; a.Tj = c->Tj; a.Tpos = c->Tpos; a.Wd = c->ee_Wd; a.bin_lower = c->bin_lower;
  a.sc_ca = sc_ca; a.ridx = ridx; a.mask = rmask;
  a.W2 = c->ee_W2; a.W3 = c->ee_W3; a.W2t = c->ee_W2t; a.W3t = c->ee_W3t;
  a.b2 = c->P(ee + "2.bias"); a.b3 = c->P(ee + "4.bias"); a.ln_w = c->P(ee + "5.weight"); a.ln_b = c->P(ee + "5.bias");
  a.z_out = z_out; a.wimg = c->ee_wimg; a.vec4 = c->ee_vec;
  // table mode (see pair_tc4.cu): when the distinct (offset, bin) rows of a decoy are well below its L^2 pair rows
  if (c->opt_pair >= 1 && c->opt_table && c->ee_table && (size_t)c->n_off * (N_BINS + 1) * 2 < (size_t)L * L &&
      edge_embed_table_elems(B, c->n_off, 4) <= c->ee_table_elems) {
    a.table_variants = 4; a.fixed = fixed; a.table = c->ee_table; a.tab_ctl = c->ee_ctl; a.cls = c->ee_cls;
  }
  // pair_kernels = 0 selects the SIMT cross-check kernels (same inputs, same rounding points); the tcgen05 kernel takes every
  // chain length the public entry points hand down (they pad to a multiple of 32)
  if (c->opt_pair >= 1) {
    S2S_CHECK(L % 32 == 0, "internal: the tcgen05 edge embedder was handed an un-padded chain length");
    edge_embed_tc2(a, st);
  } else edge_embed_simt(a, st);
}

// InvariantPointAttention.forward of block blk -> out (linear_out result; not yet masked)
void do_ipa(s2s_ctx* c, int blk, int B, int L, const float* node, const bf16* z, const float* quat, const float* trans,
            const float* rmask, float* out, const float* res, const float* row_post, cudaStream_t st, Split node_sp = Split()) {
  const int R = B * L;
  const IpaW& w = c->ipa[blk];
  const std::string ip = "translator.trunk.ipa_" + std::to_string(blk) + ".";
  const bool tc = c->opt_node == 1 && L % 16 == 0;
  const float qk_scale = 0.03608439182435161f;  // sqrt(1 / (3 * 256))   (ipa.py:187)
  // second-generation path: point term folded into the logits GEMM, tcgen05 pair kernel, split-bf16 P
  const bool fused = tc && c->opt_ipa == 1 && ipa_pair_attention_tc_supported(L);
  IpaPointsAug aug;
  if (fused) {
    aug.qp_aug = c->qp_aug; aug.kp_aug = c->kp_aug; aug.colbias = c->colbias; aug.vp_hi = c->vp_hi; aug.vp_lo = c->vp_lo;
    aug.pt_w = w.pt_w; aug.inv_alpha = 1.f / qk_scale; aug.L = L;
  }
  if (tc && !node_sp.hi) {
    split_bf16(node, 256, R, 256, c->sa_hi, c->sa_lo, st);
    node_sp.hi = c->sa_hi; node_sp.lo = c->sa_lo;
  }
  // few rows: the point projection and the rigid apply run beside the q | k | v projection (they meet at the logits GEMM)
  Fork pts(c, st, 0, fused && few_rows(R));
  if (tc) {
    // q | k | v projection in one bf16 tensor-core GEMM whose epilogue writes the attention operands directly:
    // row-major bf16 q,k (K-major for q.k^T) and the transposed v (K-major for P.v); no fp32 copy is needed.
    const auto ws = weight_split(c, w.proj_w);
    TcGemm g;
    g.A_hi = node_sp.hi; g.a_rows = R; g.a_cols = 256; g.a_pitch = 256;
    g.B_hi = ws.first; g.b_rows = 6144; g.b_cols = 256; g.b_pitch = 256;
    g.M = R; g.N = 6144; g.K = 256; g.passes = 1; g.bias = w.proj_b;
    g.out_hi = c->qkv_bf16; g.ldo = 6144;  // v stays row-major: P.v reads it as an MN-major B operand (no transposed copy)
    gemm_tc(g, st);
    // point projections keep split-bf16 accuracy (they become nm-scale coordinates)
    TcGemm p;
    p.A_hi = node_sp.hi; p.A_lo = node_sp.lo; p.a_rows = R; p.a_cols = 256; p.a_pitch = 256;
    p.B_hi = ws.first + (size_t)6144 * 256; p.B_lo = ws.second + (size_t)6144 * 256; p.b_rows = 672; p.b_cols = 256; p.b_pitch = 256;
    p.M = R; p.N = 672; p.K = 256; p.passes = 3; p.bias = w.proj_b + 6144;
    p.C = c->proj + 6144; p.ldc = 6816;
    gemm_tc(p, pts.stream());
  } else {
    linear(c, node, 256, w.proj_w, 256, w.proj_b, c->proj, 6816, R, 6816, 256, st, 0, nullptr, 0, nullptr, nullptr, EXACT);
  }
  ipa_points(c->proj + 6144, 6816, c->proj + 6336, 6816, quat, trans, c->q_pts, c->k_pts, c->v_pts, R, pts.stream(), aug);
  pts.join();
  if (tc) {  // S = scale * q.k^T, batched over (decoy, head)
    TcGemm g;
    g.A_hi = c->qkv_bf16; g.a_rows = R; g.a_cols = 6144; g.a_pitch = 6144; g.a_rb = L; g.a_ch = 256;
    g.B_hi = c->qkv_bf16 + 2048; g.b_rows = R; g.b_cols = 4096; g.b_pitch = 6144; g.b_rb = L; g.b_ch = 512;
    g.M = L; g.N = L; g.K = 256; g.nb = B; g.nh = N_H; g.passes = 1; g.alpha = qk_scale;
    g.C = c->S; g.ldc = L; g.sCb = (long)N_H * L * L; g.sCh = (long)L * L;
    if (fused) {  // + w_h q_pts.k_pts - 0.5 w_h |k_pts|^2  (= the point term up to a per-query constant)
      g.A2 = c->qp_aug; g.a2_rows = R; g.a2_cols = N_H * PT_K; g.a2_pitch = N_H * PT_K; g.a2_rb = L; g.a2_ch = PT_K;
      g.B2 = c->kp_aug; g.b2_rows = R; g.b2_cols = N_H * PT_K; g.b2_pitch = N_H * PT_K; g.b2_rb = L; g.b2_ch = PT_K;
      g.K2 = PT_K;
      g.bias = c->colbias; g.bias_sb = (long)N_H * L; g.bias_sh = L;
    }
    gemm_tc(g, st);
  } else {
    GemmArgs g;
    g.A = c->proj; g.lda = 6816; g.sAb = (long)L * 6816; g.sAh = 256;
    g.B = c->proj + 2048; g.ldb = 6816; g.sBb = (long)L * 6816; g.sBh = 512;
    g.C = c->S; g.ldc = L; g.sCb = (long)N_H * L * L; g.sCh = (long)L * L;
    g.M = L; g.N = L; g.K = 256; g.nb = B; g.nh = N_H; g.alpha = qk_scale;
    gemm_f32(g, st);
  }
  if (!fused) ipa_point_logits(c->S, c->q_pts, c->k_pts, w.pt_w, B, L, st);
  IpaPairArgs p;
  p.B = B; p.L = L; p.z = z; p.S = c->S; p.mask = rmask;
  p.Wb_hi = w.Wb_hi; p.Wb_lo = w.Wb_lo; p.bb = c->P(ip + "linear_b.bias");
  p.Wdz_t = w.Wdz_t; p.bdz = c->P(ip + "down_z.bias");
  p.o_pair = c->feats + (N_H * C_H + 4 * N_H * P_V); p.ld_opair = IPA_FEAT;
  p.P_bf16 = tc ? c->P_bf16 : nullptr;
  p.P_lo = c->P_lo; p.wb_img = w.wb_img;
  // fused path: the three producers of the 2688-wide feature row write its split-bf16 image (sa_hi / sa_lo) directly
  const long pair_col = N_H * C_H + 4 * N_H * P_V;
  if (fused) { p.opair_hi = c->sa_hi + pair_col; p.opair_lo = c->sa_lo + pair_col; }
  if (fused) ipa_pair_attention_tc(p, st); else ipa_pair_attention(p, st);
  // few rows: P.v_pts and the back-rotation of its result run beside P.v (they meet at linear_out)
  Fork vpts(c, st, 1, fused && few_rows(R));
  if (tc) {  // o = P v -> feats[:, h*256 + c]
    TcGemm g;
    g.A_hi = c->P_bf16; g.a_rows = (size_t)B * N_H * L; g.a_cols = L; g.a_pitch = L; g.a_rb = N_H * L; g.a_rh = L;
    g.B_hi = c->qkv_bf16 + 2048 + C_H; g.b_rows = R; g.b_cols = 6144 - 2048 - C_H; g.b_pitch = 6144; g.b_rb = L; g.b_ch = 2 * C_H; g.b_mn = 1;
    g.M = L; g.N = C_H; g.K = L; g.nb = B; g.nh = N_H; g.passes = 1;
    g.ldc = IPA_FEAT; g.sCb = (long)L * IPA_FEAT; g.sCh = C_H;
    if (fused) { g.out_hi = c->sa_hi; g.out_lo = c->sa_lo; g.ldo = IPA_FEAT; } else g.C = c->feats;
    gemm_tc(g, st);
  }
  if (fused) {  // o_pt (global frame) = P v_pts, split-bf16 on the tensor cores
    TcGemm g;
    g.A_hi = c->P_bf16; g.A_lo = c->P_lo; g.a_rows = (size_t)B * N_H * L; g.a_cols = L; g.a_pitch = L; g.a_rb = N_H * L; g.a_rh = L;
    g.B_hi = c->vp_hi; g.B_lo = c->vp_lo; g.b_rows = R; g.b_cols = N_H * VP_PITCH; g.b_pitch = N_H * VP_PITCH; g.b_rb = L; g.b_ch = VP_PITCH; g.b_mn = 1;  // row-major value points
    g.M = L; g.N = P_V * 3; g.K = L; g.nb = B; g.nh = N_H; g.passes = 3;
    g.C = c->opt; g.ldc = N_H * P_V * 3; g.sCb = (long)L * N_H * P_V * 3; g.sCh = P_V * 3;
    gemm_tc(g, vpts.stream());
  } else {
    GemmArgs g;
    g.A = c->S; g.lda = L; g.sAb = (long)N_H * L * L; g.sAh = (long)L * L;
    g.M = L; g.K = L; g.nb = B; g.nh = N_H; g.b_kn = 1;
    if (!tc) {  // o = P v  (exact path)
      g.B = c->proj + 2048 + 256; g.ldb = 6816; g.sBb = (long)L * 6816; g.sBh = 512;
      g.C = c->feats; g.ldc = IPA_FEAT; g.sCb = (long)L * IPA_FEAT; g.sCh = 256; g.N = 256;
      gemm_f32(g, st);
    }
    // o_pt (global frame) = P v_pts : 36 columns per head, exact fp32
    g.B = c->v_pts; g.ldb = N_H * P_V * 3; g.sBb = (long)L * N_H * P_V * 3; g.sBh = P_V * 3;
    g.C = c->opt; g.ldc = N_H * P_V * 3; g.sCb = (long)L * N_H * P_V * 3; g.sCh = P_V * 3;
    g.N = P_V * 3;
    gemm_f32(g, st);
  }
  ipa_finalize_points(c->opt, quat, trans, c->feats, R, vpts.stream(), fused ? c->sa_hi : nullptr, fused ? c->sa_lo : nullptr);
  vpts.join();
  static const int splitk_env = [] { const char* e = getenv("S2S_SPLITK"); return e ? atoi(e) : 1; }();  // 0: A/B timing
  if (fused && splitk_env && 2 * ceil_div(R, 128) <= sm_count()) {
    // few residue rows: linear_out's reduction over the 2688 features is cut into nine slices that run on different SMs
    // (c->proj, the projection buffer, is dead by now and serves as the partial-sum scratch: 9 x R x 256 of its R x 6816 floats)
    const auto w = weight_split(c, c->P(ip + "linear_out.weight"));
    TcGemm g;
    g.A_hi = c->sa_hi; g.A_lo = c->sa_lo; g.a_rows = R; g.a_cols = IPA_FEAT; g.a_pitch = IPA_FEAT;
    g.B_hi = w.first; g.B_lo = w.second; g.b_rows = 256; g.b_cols = IPA_FEAT; g.b_pitch = IPA_FEAT;
    g.M = R; g.N = 256; g.K = IPA_FEAT; g.passes = 3;
    g.bias = c->P(ip + "linear_out.bias"); g.row_post = row_post; g.res = res; g.ldres = 256; g.C = out; g.ldc = 256;
    gemm_tc_splitk(g, c->proj, st);
    return;
  }
  linear(c, c->feats, IPA_FEAT, c->P(ip + "linear_out.weight"), IPA_FEAT, c->P(ip + "linear_out.bias"), out, 256, R, 256,
         IPA_FEAT, st, 0, res, 256, nullptr, row_post, TC3, fused ? Split{c->sa_hi, c->sa_lo} : Split());
}

// prep_done: the per-residue terms (n' = initial_embed(node) with its bf16 image, u = W1[:,128:256] n' + b1, p = Wf[:,128:256] n' + bf)
// were already produced by the node-track chain of this block (do_trunk)
void do_edge_transition(s2s_ctx* c, int blk, int B, int L, const float* node, const bf16* z_in, const float* rmask,
                        bf16* z_out, cudaStream_t st, Split node_sp = Split(), bool prep_done = false) {
  const int R = B * L;
  const std::string e = "translator.trunk.edge_transition_" + std::to_string(blk) + ".";
  const float *W1 = c->P(e + "trunk.0.weight"), *Wf = c->P(e + "final_layer.weight");
  const bool tcn = c->opt_node == 1 && c->cur_prec != EXACT && L % 16 == 0;
  const Split np = tcn ? Split{c->nprime_hi, c->nprime_lo} : Split();
  const bool tc = c->opt_pair >= 1;
  S2S_CHECK(!tc || L % 32 == 0, "internal: the tcgen05 EdgeTransition was handed an un-padded chain length");
  S2S_CHECK(!prep_done || (tc && np.hi), "internal: chained EdgeTransition terms need the tensor-core paths");
  if (!prep_done) {
    linear(c, node, 256, c->P(e + "initial_embed.weight"), 256, c->P(e + "initial_embed.bias"), c->nprime, 128, R, 128, 256, st, 0,
           nullptr, 0, nullptr, nullptr, -1, node_sp, np);
    Fork pf(c, st, 3, np.hi != nullptr && few_rows(R));  // few rows: the two per-residue terms side by side
    linear(c, c->nprime, 128, Wf + 128, 384, c->P(e + "final_layer.bias"), c->p128, 128, R, 128, 128, pf.stream(), 0, nullptr, 0, nullptr, nullptr, -1, np);
    linear(c, c->nprime, 128, W1 + 128, 384, c->P(e + "trunk.0.bias"), c->u384, 384, R, 384, 128, st, 0, nullptr, 0, nullptr, nullptr, -1, np);
    pf.join();
  }
  if (tc) {  // the n'_j terms ride along as extra K columns of the MMAs: they need n' in bf16 (= the hi image)
    if (!np.hi) f32_to_bf16(c->nprime, c->nprime_bf16, (long)R * 128, st);
  } else {
    linear(c, c->nprime, 128, W1 + 256, 384, nullptr, c->v384, 384, R, 384, 128, st, 0, nullptr, 0, nullptr, nullptr, -1, np);
    linear(c, c->nprime, 128, Wf + 256, 384, nullptr, c->q128, 128, R, 128, 128, st, 0, nullptr, 0, nullptr, nullptr, -1, np);
  }
  EdgeTransitionArgs a;
  a.B = B; a.L = L; a.z_in = z_in; a.u = c->u384; a.v = c->v384; a.p = c->p128; a.q = c->q128; a.mask = rmask;
  const EtW& w = c->et[blk];
  a.W1z = w.W1z; a.W2 = w.W2; a.Wfh = w.Wfh; a.Wfz = w.Wfz;
  a.W1zt = w.W1zt; a.W2t = w.W2t; a.Wft = w.Wft; a.Wfzt = w.Wfzt;
  a.b2 = c->P(e + "trunk.2.bias"); a.ln_w = c->P(e + "layer_norm.weight"); a.ln_b = c->P(e + "layer_norm.bias");
  a.z_out = z_out; a.wimg3 = w.wimg3; a.wimg_copies = c->wimg_copies; a.nprime_bf16 = c->nprime_bf16;
  if (!tc) edge_transition_simt(a, st);
  else if (c->opt_et_pair && edge_transition_pair_supported(B, L)) edge_transition_pair(a, st);
  else edge_transition_tc3(a, st);
}

// self-attention of one sequence-transformer layer on c->x320 -> c->y320 (in_proj, q.k^T, key-biased softmax, P.v);
// in_proj_done: the in_proj GEMM was the last step of the previous chain launch (c->tq_hi already holds q | k | v)
void do_tfm_attention(s2s_ctx* c, const std::string& tl, int B, int L, cudaStream_t st, bool in_proj_done = false) {
  const int R = B * L;
  const float scale = 0.11180339887498948f;  // 1/sqrt(80)
  const bool tc = c->opt_node == 1 && L % 16 == 0;
  if (tc) {
    // in_proj on the tensor cores; its epilogue emits the split-bf16 attention operands: q|k row-major, v transposed per head
    const auto w = weight_split(c, c->P(tl + "self_attn.in_proj_weight"));
    TcGemm g;  // A = x320's split image, written by the producer (concat / norm2 of the previous layer)
    const bool sp = c->tfm_passes == 3;  // single-pass attention (measured: no change of the trajectory error) needs no lo images
    // in_proj itself stays split-bf16: single pass costs 2.4x of the trajectory error budget (3.0e-5 vs 1.27e-5) for 0.14 ms
    g.A_hi = c->x320_hi; g.A_lo = c->x320_lo; g.a_rows = R; g.a_cols = 320; g.a_pitch = 320;
    g.B_hi = w.first; g.B_lo = w.second; g.b_rows = 960; g.b_cols = 320; g.b_pitch = 320;
    g.M = R; g.N = 960; g.K = 320; g.passes = 3; g.bias = c->P(tl + "self_attn.in_proj_bias");
    g.out_hi = c->tq_hi; g.out_lo = sp ? c->tq_lo : nullptr; g.ldo = 960;
    if (!in_proj_done) gemm_tc(g, st);
    static const int fused_attn = [] { const char* e = getenv("S2S_TFM_FUSED"); return e ? atoi(e) : 1; }();
    if (fused_attn && !sp && tfm_attention_supported(L)) {
      // q.k^T, key-biased softmax and P.v in one kernel: logits and weights never leave the SM
      tfm_attention(c->tq_hi, c->keybias, c->y320, c->y320_hi, c->y320_lo, B, L, scale, st);
    } else {
    TcGemm s;  // logits = q.k^T / sqrt(80), batched over (decoy, head), split-bf16
    s.A_hi = c->tq_hi; s.A_lo = c->tq_lo; s.a_rows = R; s.a_cols = 960; s.a_pitch = 960; s.a_rb = L; s.a_ch = TFM_HD;
    s.B_hi = c->tq_hi + 320; s.B_lo = c->tq_lo + 320; s.b_rows = R; s.b_cols = 640; s.b_pitch = 960; s.b_rb = L; s.b_ch = TFM_HD;
    s.M = L; s.N = L; s.K = TFM_HD; s.nb = B; s.nh = TFM_H; s.passes = c->tfm_passes; s.alpha = scale;
    s.C = c->S; s.ldc = L; s.sCb = (long)TFM_H * L * L; s.sCh = (long)L * L;
    gemm_tc(s, st);
    softmax_keybias(c->S, c->keybias, B, TFM_H, L, st, c->P_bf16, sp ? c->tP_lo : nullptr);
    TcGemm p;  // y = P v
    p.A_hi = c->P_bf16; p.A_lo = c->tP_lo; p.a_rows = (size_t)B * TFM_H * L; p.a_cols = L; p.a_pitch = L; p.a_rb = TFM_H * L; p.a_rh = L;
    p.B_hi = c->tq_hi + 640; p.B_lo = c->tq_lo + 640; p.b_rows = R; p.b_cols = 320; p.b_pitch = 960; p.b_rb = L; p.b_ch = TFM_HD; p.b_mn = 1;  // v columns, read MN-major
    p.M = L; p.N = TFM_HD; p.K = L; p.nb = B; p.nh = TFM_H; p.passes = c->tfm_passes;
    p.C = c->y320; p.ldc = 320; p.sCb = (long)L * 320; p.sCh = TFM_HD;
    p.out_hi = c->y320_hi; p.out_lo = c->y320_lo; p.ldo = 320;
    gemm_tc(p, st);
    }
  } else {
    linear(c, c->x320, 320, c->P(tl + "self_attn.in_proj_weight"), 320, c->P(tl + "self_attn.in_proj_bias"), c->qkv, 960, R, 960, 320, st);
    GemmArgs g;
    g.A = c->qkv; g.lda = 960; g.sAb = (long)L * 960; g.sAh = TFM_HD;
    g.B = c->qkv + 320; g.ldb = 960; g.sBb = (long)L * 960; g.sBh = TFM_HD;
    g.C = c->S; g.ldc = L; g.sCb = (long)TFM_H * L * L; g.sCh = (long)L * L;
    g.M = L; g.N = L; g.K = TFM_HD; g.nb = B; g.nh = TFM_H; g.alpha = scale;
    gemm_f32(g, st);
    softmax_keybias(c->S, c->keybias, B, TFM_H, L, st);
    GemmArgs h;
    h.A = c->S; h.lda = L; h.sAb = (long)TFM_H * L * L; h.sAh = (long)L * L;
    h.B = c->qkv + 640; h.ldb = 960; h.sBb = (long)L * 960; h.sBh = TFM_HD; h.b_kn = 1;
    h.C = c->y320; h.ldc = 320; h.sCb = (long)L * 320; h.sCh = TFM_HD;
    h.M = L; h.N = TFM_HD; h.K = L; h.nb = B; h.nh = TFM_H;
    gemm_f32(h, st);
  }
}

// the row-local rest of the layer (post-norm nn.TransformerEncoderLayer): out_proj + residual + norm1, linear1 + ReLU,
// linear2 + residual + norm2, one launch per layer / LayerNorm
void do_tfm_tail(s2s_ctx* c, const std::string& tl, int B, int L, cudaStream_t st) {
  const int R = B * L;
  const bool tc = c->opt_node == 1 && L % 16 == 0;
  const Split x_sp = tc ? Split{c->x320_hi, c->x320_lo} : Split();
  const Split t_sp = tc ? Split{c->t320_hi, c->t320_lo} : Split();
  const Split y_sp = tc ? Split{c->y320_hi, c->y320_lo} : Split();
  linear(c, c->y320, 320, c->P(tl + "self_attn.out_proj.weight"), 320, c->P(tl + "self_attn.out_proj.bias"), c->t320, 320, R, 320, 320, st, 0, c->x320, 320,
         nullptr, nullptr, -1, y_sp);
  layernorm(c->t320, nullptr, c->P(tl + "norm1.weight"), c->P(tl + "norm1.bias"), nullptr, c->x320, R, 320, st, x_sp.hi, x_sp.lo);
  linear(c, c->x320, 320, c->P(tl + "linear1.weight"), 320, c->P(tl + "linear1.bias"), c->t320, 320, R, 320, 320, st, 1, nullptr, 0, nullptr, nullptr, -1, x_sp, t_sp);
  linear(c, c->t320, 320, c->P(tl + "linear2.weight"), 320, c->P(tl + "linear2.bias"), c->y320, 320, R, 320, 320, st, 0, c->x320, 320, nullptr, nullptr, -1, t_sp);
  layernorm(c->y320, nullptr, c->P(tl + "norm2.weight"), c->P(tl + "norm2.bias"), nullptr, c->x320, R, 320, st, x_sp.hi, x_sp.lo);
}


// ---- gemm_chain steps (row-local layers fused into one launch; see gemm_chain.cu) ----
ChainStep chain_step(const s2s_ctx* c, Split in, int K, const float* W, long ldw, const float* bias, int N, int relu = 0,
                     float* C = nullptr, long ldc = 0, const float* res = nullptr, long ldres = 0, Split out = Split()) {
  const auto w = weight_split(c, W);
  ChainStep t;
  t.A_hi = in.hi; t.A_lo = in.lo; t.W_hi = w.first; t.W_lo = w.second; t.ldw = ldw;
  t.N = N; t.K = K; t.relu = relu; t.bias = bias; t.res = res; t.ldres = ldres; t.C = C; t.ldc = ldc;
  t.out_hi = out.hi; t.out_lo = out.lo; t.ldo = N;
  return t;
}
void chain_ln(ChainStep& t, const float* w, const float* b, const float* scale, float* out, Split img) {
  t.ln_w = w; t.ln_b = b; t.ln_scale = scale; t.ln_out = out; t.ln_hi = img.hi; t.ln_lo = img.lo; t.ld_ln = t.N;
}
// do_tfm_tail as chain steps: y320 image -> x320 (+ image)
void tfm_tail_steps(const s2s_ctx* c, const std::string& tl, std::vector<ChainStep>& v) {
  const Split x_sp{c->x320_hi, c->x320_lo}, t_sp{c->t320_hi, c->t320_lo}, y_sp{c->y320_hi, c->y320_lo};
  ChainStep o = chain_step(c, y_sp, 320, c->P(tl + "self_attn.out_proj.weight"), 320, c->P(tl + "self_attn.out_proj.bias"), 320, 0, c->t320, 320, c->x320, 320);
  chain_ln(o, c->P(tl + "norm1.weight"), c->P(tl + "norm1.bias"), nullptr, c->x320, x_sp);
  v.push_back(o);
  v.push_back(chain_step(c, x_sp, 320, c->P(tl + "linear1.weight"), 320, c->P(tl + "linear1.bias"), 320, 1, nullptr, 0, nullptr, 0, t_sp));
  ChainStep l2 = chain_step(c, t_sp, 320, c->P(tl + "linear2.weight"), 320, c->P(tl + "linear2.bias"), 320, 0, c->y320, 320, c->x320, 320);
  chain_ln(l2, c->P(tl + "norm2.weight"), c->P(tl + "norm2.bias"), nullptr, c->x320, x_sp);
  v.push_back(l2);
}

// TranslationIPA.forward (ipa.py:331-387) on the node / pair embeddings already in c->node / c->z
void do_trunk(s2s_ctx* c, int B, int L, const float* rigids_t, const float* rmask, const float* fixed,
              const float* gt_psi, float* out_rigids, float* out_psi, cudaStream_t st) {
  const int R = B * L;
  const std::string tk = "translator.trunk.";
  make_masks(rmask, fixed, c->hard, c->diffuse, c->keybias, R, st);
  PrecScope prec_scope(c, TC3);  // residual-stream layers: split-bf16 tensor-core GEMMs (exact fp32 when node_gemm = 0)
  // Split-bf16 images of the activations are written by whichever kernel produces them (LayerNorm, GEMM epilogue,
  // concat), so the tensor-core GEMMs that consume them need no separate conversion pass.
  const bool img = c->opt_node == 1 && L % 16 == 0;
  // chain: the row-local layers between the attention kernels run as gemm_chain launches (same arithmetic, same bits).
  // Measured on B200 (bench.py, CUDA-graph replay with programmatic dependent launch): 49.27 -> 49.44 and 49.11 -> 49.26
  // conformations/s at cfg2 (16384 residue rows, 162 -> 88 launches per forward), but 130.4 -> 122.1 at L = 64 x 32 decoys and
  // 46.7 -> 44.2 at L = 64 x 1: with fewer row panels than SMs a layer is bound by its own load -> stage -> MMA -> epilogue
  // latency chain, which a step of the chain does not shorten, while separate launches overlap their prologues on idle SMs.
  // So option "chain" = 1 (default) chains when there are more 128-row panels than half the SMs — below that the separate panel
  // GEMMs split their output columns over the idle SMs instead (gemm_tc.cu: n_split) —, 2 always (tests), 0 never.
  const bool chain = img && (c->opt_chain == 2 || (c->opt_chain == 1 && 2 * ceil_div(R, 128) > sm_count()));
  auto sp = [&](bf16* hi, bf16* lo) { return img ? Split{hi, lo} : Split(); };
  const Split node_sp = sp(c->node_hi, c->node_lo), init_sp = sp(c->init_hi, c->init_lo), a_sp = sp(c->a256_hi, c->a256_lo),
              b_sp = sp(c->b256_hi, c->b256_lo), x_sp = sp(c->x320_hi, c->x320_lo);
  if (img) {
    // every block's skip connection reads the same initial node embedding (ipa.py:333,353-356): one stacked GEMM serves the
    // four of them, here, while c->node still IS that embedding (so no copy of it is kept)
    split_bf16(c->node, 256, R, 256, c->node_hi, c->node_lo, st);
    linear(c, c->node, 256, c->skip_w_all, 256, c->skip_b_all, c->skip_all, N_BLK * D_SKIP, R, N_BLK * D_SKIP, 256, st, 0, nullptr, 0, nullptr, nullptr, -1, node_sp);
  } else {
    S2S_CUDA(cudaMemcpyAsync(c->init_node, c->node, (size_t)R * 256 * 4, cudaMemcpyDeviceToDevice, st));
  }
  split_rigids(rigids_t, c->quat, c->trans, R, st);
  const std::string tp = "translator.torsion_pred.";
  for (int b = 0; b < N_BLK; ++b) {
    const std::string s = std::to_string(b);
    const std::string nt = tk + "node_transition_" + s + ".", t0 = tk + "transformer_" + s + ".layers.0.", t1 = tk + "transformer_" + s + ".layers.1.";
    // node = LN(node + ipa(node, z, T) * mask)          (ipa.py:344-351)
    do_ipa(c, b, B, L, c->node, c->z, c->quat, c->trans, rmask, c->a256, c->node, rmask, st, node_sp);
    if (img) {
      // ... and the sequence transformer's input [node | skip(init_node)] (ipa.py:353-356) from the same LayerNorm launch
      LnExtra ex;
      ex.ld2 = D_TFM; ex.y2 = c->x320; ex.tail = c->skip_all + b * D_SKIP; ex.tail_ld = N_BLK * D_SKIP; ex.tail_w = D_SKIP;
      layernorm(c->a256, nullptr, c->P(tk + "ipa_ln_" + s + ".weight"), c->P(tk + "ipa_ln_" + s + ".bias"), nullptr, c->node, R, 256, st, x_sp.hi, x_sp.lo, ex);
    } else {
      layernorm(c->a256, nullptr, c->P(tk + "ipa_ln_" + s + ".weight"), c->P(tk + "ipa_ln_" + s + ".bias"), nullptr, c->node, R, 256, st);
      // sequence transformer on [node | skip(init_node)]   (ipa.py:353-360)
      linear(c, c->init_node, 256, c->P(tk + "skip_embed_" + s + ".weight"), 256, c->P(tk + "skip_embed_" + s + ".bias"), c->skip64, 64, R, 64, 256, st, 0,
             nullptr, 0, nullptr, nullptr, -1, init_sp);
      concat_skip(c->node, c->skip64, c->x320, R, st, x_sp.hi, x_sp.lo);
    }
    if (chain) {
      const bool sp3 = c->tfm_passes == 3;
      std::vector<ChainStep> v;
      // layer 0: attention, then its row-local tail and layer 1's in_proj in one launch
      do_tfm_attention(c, t0, B, L, st);
      tfm_tail_steps(c, t0, v);
      v.push_back(chain_step(c, x_sp, 320, c->P(t1 + "self_attn.in_proj_weight"), 320, c->P(t1 + "self_attn.in_proj_bias"), 960, 0, nullptr, 0, nullptr, 0,
                             Split{c->tq_hi, sp3 ? c->tq_lo : nullptr}));
      gemm_chain(v.data(), (int)v.size(), R, st);
      v.clear();
      // layer 1: attention, then everything up to the next kernel that looks across residues
      do_tfm_attention(c, t1, B, L, st, true);
      tfm_tail_steps(c, t1, v);
      // node += linear(transformer output)                 (ipa.py:359-360)
      v.push_back(chain_step(c, x_sp, 320, c->P(tk + "linear_" + s + ".weight"), 320, c->P(tk + "linear_" + s + ".bias"), 256, 0, c->node, 256, c->node, 256, node_sp));
      // node transition + mask                              (layers.py:138-145, ipa.py:363-365)
      v.push_back(chain_step(c, node_sp, 256, c->P(nt + "linear_1.weight"), 256, c->P(nt + "linear_1.bias"), 256, 1, nullptr, 0, nullptr, 0, a_sp));
      v.push_back(chain_step(c, a_sp, 256, c->P(nt + "linear_2.weight"), 256, c->P(nt + "linear_2.bias"), 256, 1, nullptr, 0, nullptr, 0, b_sp));
      ChainStep l3 = chain_step(c, b_sp, 256, c->P(nt + "linear_3.weight"), 256, c->P(nt + "linear_3.bias"), 256, 0, c->a256, 256, c->node, 256);
      chain_ln(l3, c->P(nt + "ln.weight"), c->P(nt + "ln.bias"), rmask, c->node, node_sp);
      v.push_back(l3);
      if (b < N_BLK - 1 && c->opt_pair >= 1) {
        // per-residue terms of the EdgeTransition (layers.py:170-176): n' = initial_embed(node), u = W1[:,128:256] n' + b1, p = Wf[:,128:256] n' + bf
        const std::string e = tk + "edge_transition_" + s + ".";
        const Split np{c->nprime_hi, c->nprime_lo};
        v.push_back(chain_step(c, node_sp, 256, c->P(e + "initial_embed.weight"), 256, c->P(e + "initial_embed.bias"), 128, 0, c->nprime, 128, nullptr, 0, np));
        v.push_back(chain_step(c, np, 128, c->P(e + "trunk.0.weight") + 128, 384, c->P(e + "trunk.0.bias"), 384, 0, c->u384, 384));
        v.push_back(chain_step(c, np, 128, c->P(e + "final_layer.weight") + 128, 384, c->P(e + "final_layer.bias"), 128, 0, c->p128, 128));
      } else if (b == N_BLK - 1) {
        // torsion head (layers.py:199-213) up to its exact-fp32 last layer
        v.push_back(chain_step(c, node_sp, 256, c->P(tp + "linear_1.weight"), 256, c->P(tp + "linear_1.bias"), 256, 1, nullptr, 0, nullptr, 0, a_sp));
        v.push_back(chain_step(c, a_sp, 256, c->P(tp + "linear_2.weight"), 256, c->P(tp + "linear_2.bias"), 256, 0, c->b256, 256, c->node, 256));
      }
      gemm_chain(v.data(), (int)v.size(), R, st);
    } else {
      do_tfm_attention(c, t0, B, L, st);
      do_tfm_tail(c, t0, B, L, st);
      do_tfm_attention(c, t1, B, L, st);
      do_tfm_tail(c, t1, B, L, st);
      linear(c, c->x320, 320, c->P(tk + "linear_" + s + ".weight"), 320, c->P(tk + "linear_" + s + ".bias"), c->node, 256, R, 256, 320, st, 0, c->node, 256,
             nullptr, nullptr, -1, x_sp, node_sp);
      // node transition + mask                              (layers.py:138-145, ipa.py:363-365)
      linear(c, c->node, 256, c->P(nt + "linear_1.weight"), 256, c->P(nt + "linear_1.bias"), c->a256, 256, R, 256, 256, st, 1, nullptr, 0, nullptr, nullptr, -1, node_sp, a_sp);
      linear(c, c->a256, 256, c->P(nt + "linear_2.weight"), 256, c->P(nt + "linear_2.bias"), c->b256, 256, R, 256, 256, st, 1, nullptr, 0, nullptr, nullptr, -1, a_sp, b_sp);
      linear(c, c->b256, 256, c->P(nt + "linear_3.weight"), 256, c->P(nt + "linear_3.bias"), c->a256, 256, R, 256, 256, st, 0, c->node, 256, nullptr, nullptr, -1, b_sp);
      layernorm(c->a256, nullptr, c->P(nt + "ln.weight"), c->P(nt + "ln.bias"), rmask, c->node, R, 256, st, node_sp.hi, node_sp.lo);
    }
    // backbone update on node * diffuse_mask, exact fp32    (ipa.py:367-369)
    // (few rows: beside the EdgeTransition's per-residue GEMMs, which read the same node rows and nothing the update writes)
    Fork bb(c, st, 2, img && few_rows(R) && b < N_BLK - 1);
    bb_update_frame(c->node, c->P(tk + "bb_update_" + s + ".linear.weight"), c->P(tk + "bb_update_" + s + ".linear.bias"), c->quat, c->trans, c->diffuse, R, bb.stream());
    if (b < N_BLK - 1) do_edge_transition(c, b, B, L, c->node, c->z, rmask, c->z, st, node_sp, chain && c->opt_pair >= 1);
    bb.join();
  }
  if (!chain) {
    linear(c, c->node, 256, c->P(tp + "linear_1.weight"), 256, c->P(tp + "linear_1.bias"), c->a256, 256, R, 256, 256, st, 1, nullptr, 0, nullptr, nullptr, -1, node_sp, a_sp);
    linear(c, c->a256, 256, c->P(tp + "linear_2.weight"), 256, c->P(tp + "linear_2.bias"), c->b256, 256, R, 256, 256, st, 0, c->node, 256, nullptr, nullptr, -1, a_sp);
  }
  linear(c, c->b256, 256, c->P(tp + "linear_final.weight"), 256, c->P(tp + "linear_final.bias"), c->psi_u, 2, R, 2, 256, st, 0, nullptr, 0, nullptr, nullptr, EXACT);
  c->cur_prec = EXACT;
  psi_finalize(c->psi_u, gt_psi, fixed, out_psi, R, st);
  join_rigids(c->quat, c->trans, out_rigids, R, st);
}

// padded copies of the per-residue inputs (any of them may be null) into the context's pd_* buffers
void pad_residue_inputs(s2s_ctx* c, int B, int L, int Lp, const float* rig, const float* sc, const long long* ridx, const float* rmask,
                        const float* fixed, const float* psi, cudaStream_t st) {
  PadInputs p;
  p.B = B; p.L = L; p.Lp = Lp;
  p.rig = rig; p.sc = sc; p.ridx = ridx; p.rmask = rmask; p.fixed = fixed; p.psi = psi;
  p.o_rig = c->pd_rig; p.o_sc = c->pd_sc; p.o_ridx = c->pd_ridx; p.o_rmask = c->pd_rmask; p.o_fixed = c->pd_fixed; p.o_psi = c->pd_psi;
  p.o_hard = c->pd_hard;
  pad_inputs(p, st);
}

void do_forward(s2s_ctx* c, int B, int L, const float* rigids_t, const float* sc_ca, const float* t,
                const long long* ridx, const float* rmask, const float* fixed, const float* gt_psi, float* out_rigids,
                float* out_psi, cudaStream_t st) {
  check_shape(c, B, L);
  const int Lp = pad_len(L);
  if (Lp == L) {
    do_embed(c, B, L, t, ridx, fixed, sc_ca, rmask, c->node, c->z, st);
    do_trunk(c, B, L, rigids_t, rmask, fixed, gt_psi, out_rigids, out_psi, st);
    return;
  }
  pad_residue_inputs(c, B, L, Lp, rigids_t, sc_ca, ridx, rmask, fixed, gt_psi, st);
  HardScope hs(c, c->pd_hard);
  do_embed(c, B, Lp, t, c->pd_ridx, c->pd_fixed, c->pd_sc, c->pd_rmask, c->node, c->z, st);
  do_trunk(c, B, Lp, c->pd_rig, c->pd_rmask, c->pd_fixed, c->pd_psi, c->pd_orig, c->pd_opsi, st);
  repitch_rows(c->pd_orig, out_rigids, B, Lp, L, 7, st);
  repitch_rows(c->pd_opsi, out_psi, B, Lp, L, 2, st);
}

// ---- module-level pieces of the trunk (reference layers.py:138-145,199-213,232-241) on `rows` residue rows ----
void check_rows(const s2s_ctx* c, long rows) {
  S2S_CHECK(c->finalized, "context not finalized");
  S2S_CHECK(rows > 0 && c->ws.base && rows <= (long)c->cap_B * c->cap_L, "workspace too small: call s2s_reserve first");
}

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

int s2s_abi_version(void) { return S2S_ABI_VERSION; }
void s2s_profile_enable(int on) { g_profile_on = on != 0; }
void s2s_profile_reset(void) { prof_drain(); g_prof.clear(); }
int s2s_profile_list(char* buf, int cap) {
  prof_drain();
  std::string out;
  for (auto& kv : g_prof) out += kv.first + "\t" + std::to_string(kv.second.ms) + "\t" + std::to_string(kv.second.n) + "\n";
  if ((int)out.size() + 1 > cap) return -(int)out.size() - 1;
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int)out.size();
}
int s2s_profile_read(const char* name, double* total_ms, int64_t* count) {
  prof_drain();
  auto it = g_prof.find(name ? name : "");
  if (it == g_prof.end()) return 1;
  if (total_ms) *total_ms = it->second.ms;
  if (count) *count = it->second.n;
  return 0;
}
const char* s2s_last_error(void) { return g_error.c_str(); }
int64_t s2s_launch_count(void) { return g_launch_count; }

s2s_ctx* s2s_create(const float* tfreq, const float* pdenom, const float* bin_lower, const float* backbone) {
  s2s_ctx* c = nullptr;
  const int rc = guarded([&] {
    S2S_CHECK(tfreq && pdenom && bin_lower && backbone, "s2s_create: null table");
    c = new s2s_ctx();
    if (const char* e = getenv("S2S_WIMG_COPIES")) c->wimg_copies = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("S2S_TFM_PASSES")) c->tfm_passes = atoi(e) == 3 ? 3 : 1;
    if (const char* e = getenv("S2S_ET_PAIR")) c->opt_et_pair = atoi(e) != 0;
    if (const char* e = getenv("S2S_CHAIN")) c->opt_chain = std::max(0, std::min(2, atoi(e)));  // A/B timing; s2s_set_option("chain") overrides
    if (const char* e = getenv("S2S_FORK")) c->opt_fork = atoi(e) != 0;
    S2S_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) {
      S2S_CUDA(cudaEventCreateWithFlags(&c->ev_fork[k], cudaEventDisableTiming));
      S2S_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    auto up = [&](const float* h, size_t n) {
      float* d;
      S2S_CUDA(cudaMalloc(&d, n * 4));
      S2S_CUDA(cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice));
      return d;
    };
    c->tfreq = up(tfreq, 16); c->pdenom = up(pdenom, 16); c->bin_lower = up(bin_lower, N_BINS); c->backbone = up(backbone, 21 * 33);
  });
  if (rc) { delete c; return nullptr; }
  return c;
}

void s2s_destroy(s2s_ctx* c) {
  if (!c) return;
  c->ws.release(); c->wslab.release();
  if (c->side) cudaStreamDestroy(c->side);
  for (int k = 0; k < 4; ++k) {
    if (c->ev_fork[k]) cudaEventDestroy(c->ev_fork[k]);
    if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
  }
  cudaFree(c->tfreq); cudaFree(c->pdenom); cudaFree(c->bin_lower); cudaFree(c->backbone);
  delete c;
}

int s2s_set_param(s2s_ctx* c, const char* name, const float* data, int64_t numel) {
  return guarded([&] {
    S2S_CHECK(c && name && data && numel > 0, "s2s_set_param: bad argument");
    c->params[name] = {data, numel};
    c->finalized = false;
  });
}

int s2s_finalize(s2s_ctx* c, void* stream) {
  return guarded([&] { S2S_CHECK(c, "null ctx"); do_finalize(c, (cudaStream_t)stream); });
}

int s2s_set_option(s2s_ctx* c, const char* key, int value) {
  return guarded([&] {
    S2S_CHECK(c && key, "null argument");
    const std::string k = key;
    if (k == "pair_kernels") { S2S_CHECK(value == 0 || value == 1, "pair_kernels: 0|1"); c->opt_pair = value; }
    else if (k == "node_gemm") { S2S_CHECK(value == 0 || value == 1, "node_gemm: 0|1"); c->opt_node = value; }
    else if (k == "ipa_kernels") { S2S_CHECK(value == 0 || value == 1, "ipa_kernels: 0|1"); c->opt_ipa = value; }
    else if (k == "et_pair") { S2S_CHECK(value == 0 || value == 1, "et_pair: 0|1"); c->opt_et_pair = value; }
    else if (k == "chain") { S2S_CHECK(value >= 0 && value <= 2, "chain: 0|1|2"); c->opt_chain = value; }
    else if (k == "embed_table") { S2S_CHECK(value == 0 || value == 1, "embed_table: 0|1"); c->opt_table = value; }
    else S2S_CHECK(false, "unknown option " + k);
  });
}

int s2s_reserve(s2s_ctx* c, int B, int L, int d_min, int d_max, void* stream) {
  return guarded([&] { S2S_CHECK(c, "null ctx"); do_reserve(c, B, L, d_min, d_max, (cudaStream_t)stream); });
}

int s2s_net_forward(s2s_ctx* c, int B, int L, const float* rigids_t, const float* sc_ca, const float* t,
                    const int64_t* residue_idx, const float* residue_mask, const float* fixed_mask, const float* gt_psi,
                    float* out_rigids, float* out_psi, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && rigids_t && sc_ca && t && residue_idx && residue_mask && fixed_mask && gt_psi && out_rigids && out_psi, "s2s_net_forward: null argument");
    do_forward(c, B, L, rigids_t, sc_ca, t, (const long long*)residue_idx, residue_mask, fixed_mask, gt_psi, out_rigids, out_psi, (cudaStream_t)stream);
  });
}

int s2s_trunk(s2s_ctx* c, int B, int L, const float* node_embed, const void* z, const float* rigids_t,
              const float* residue_mask, const float* fixed_mask, const float* gt_psi, float* out_rigids, float* out_psi,
              void* stream) {
  return guarded([&] {
    S2S_CHECK(c && node_embed && z && rigids_t && residue_mask && fixed_mask && out_rigids && out_psi, "s2s_trunk: null argument");
    check_shape(c, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t R = (size_t)B * L;
    const int Lp = pad_len(L);
    if (Lp == L) {
      if (node_embed != c->node) S2S_CUDA(cudaMemcpyAsync(c->node, node_embed, R * 256 * 4, cudaMemcpyDeviceToDevice, st));
      if (z != (const void*)c->z) S2S_CUDA(cudaMemcpyAsync(c->z, z, R * L * C_Z * 2, cudaMemcpyDeviceToDevice, st));
      do_trunk(c, B, L, rigids_t, residue_mask, fixed_mask, gt_psi, out_rigids, out_psi, st);
      return;
    }
    S2S_CHECK(node_embed != c->node && z != (const void*)c->z, "s2s_trunk: padded chain lengths need caller-owned embeddings");
    repitch_rows(node_embed, c->node, B, L, Lp, 256, st);
    repitch_pair((const bf16*)z, c->z, B, L, Lp, st);
    pad_residue_inputs(c, B, L, Lp, rigids_t, nullptr, nullptr, residue_mask, fixed_mask, gt_psi, st);
    HardScope hs(c, c->pd_hard);
    do_trunk(c, B, Lp, c->pd_rig, c->pd_rmask, c->pd_fixed, gt_psi ? c->pd_psi : nullptr, c->pd_orig, c->pd_opsi, st);
    repitch_rows(c->pd_orig, out_rigids, B, Lp, L, 7, st);
    repitch_rows(c->pd_opsi, out_psi, B, Lp, L, 2, st);
  });
}

int s2s_embed(s2s_ctx* c, int B, int L, const float* t, const int64_t* residue_idx, const float* fixed_mask,
              const float* sc_ca, const float* residue_mask, float* node_out, void* z_out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && t && residue_idx && fixed_mask && sc_ca && residue_mask && node_out && z_out, "s2s_embed: null argument");
    check_shape(c, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = pad_len(L);
    if (Lp == L) {
      do_embed(c, B, L, t, (const long long*)residue_idx, fixed_mask, sc_ca, residue_mask, node_out, (bf16*)z_out, st);
      return;
    }
    pad_residue_inputs(c, B, L, Lp, nullptr, sc_ca, (const long long*)residue_idx, residue_mask, fixed_mask, nullptr, st);
    HardScope hs(c, c->pd_hard);
    do_embed(c, B, Lp, t, c->pd_ridx, c->pd_fixed, c->pd_sc, c->pd_rmask, c->node, c->z, st);
    repitch_rows(c->node, node_out, B, Lp, L, 256, st);
    repitch_pair(c->z, (bf16*)z_out, B, Lp, L, st);
  });
}

int s2s_ipa(s2s_ctx* c, int blk, int B, int L, const float* node, const void* z, const float* quat, const float* trans_nm,
            const float* residue_mask, float* out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && node && z && quat && trans_nm && residue_mask && out && blk >= 0 && blk < N_BLK, "s2s_ipa: bad argument");
    check_shape(c, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = pad_len(L);
    if (Lp == L) {
      do_ipa(c, blk, B, L, node, (const bf16*)z, quat, trans_nm, residue_mask, out, nullptr, nullptr, st);
      return;
    }
    repitch_rows(node, c->node, B, L, Lp, 256, st);
    repitch_pair((const bf16*)z, c->z, B, L, Lp, st);
    repitch_rows(quat, c->quat, B, L, Lp, 4, st);      // appended rows: zero quaternion / translation (finite, masked out)
    repitch_rows(trans_nm, c->trans, B, L, Lp, 3, st);
    pad_residue_inputs(c, B, L, Lp, nullptr, nullptr, nullptr, residue_mask, nullptr, nullptr, st);
    HardScope hs(c, c->pd_hard);
    do_ipa(c, blk, B, Lp, c->node, c->z, c->quat, c->trans, c->pd_rmask, c->a256, nullptr, nullptr, st);
    repitch_rows(c->a256, out, B, Lp, L, 256, st);
  });
}

int s2s_edge_transition(s2s_ctx* c, int blk, int B, int L, const float* node, const void* z_in, const float* residue_mask,
                        void* z_out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && node && z_in && residue_mask && z_out && blk >= 0 && blk < N_BLK - 1, "s2s_edge_transition: bad argument");
    check_shape(c, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = pad_len(L);
    PrecScope ps(c, TC3);  // n' and the per-residue terms: split-bf16 tensor-core GEMMs as inside the trunk (exact when node_gemm = 0)
    if (Lp == L) {
      do_edge_transition(c, blk, B, L, node, (const bf16*)z_in, residue_mask, (bf16*)z_out, st);
      return;
    }
    repitch_rows(node, c->node, B, L, Lp, 256, st);
    repitch_pair((const bf16*)z_in, c->z, B, L, Lp, st);
    pad_residue_inputs(c, B, L, Lp, nullptr, nullptr, nullptr, residue_mask, nullptr, nullptr, st);
    HardScope hs(c, c->pd_hard);
    do_edge_transition(c, blk, B, Lp, c->node, c->z, c->pd_rmask, c->z, st);
    repitch_pair(c->z, (bf16*)z_out, B, Lp, L, st);
  });
}

// NodeTransition.forward of block blk (layers.py:138-145): out = LN(s + linear_3(relu(linear_2(relu(linear_1(s))))))
int s2s_node_transition(s2s_ctx* c, int blk, int64_t rows, const float* s_in, float* out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && s_in && out && blk >= 0 && blk < N_BLK, "s2s_node_transition: bad argument");
    check_rows(c, rows);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = (int)rows;
    const std::string nt = "translator.trunk.node_transition_" + std::to_string(blk) + ".";
    PrecScope ps(c, TC3);
    linear(c, s_in, 256, c->P(nt + "linear_1.weight"), 256, c->P(nt + "linear_1.bias"), c->a256, 256, R, 256, 256, st, 1);
    linear(c, c->a256, 256, c->P(nt + "linear_2.weight"), 256, c->P(nt + "linear_2.bias"), c->b256, 256, R, 256, 256, st, 1);
    linear(c, c->b256, 256, c->P(nt + "linear_3.weight"), 256, c->P(nt + "linear_3.bias"), c->a256, 256, R, 256, 256, st, 0, s_in, 256);
    layernorm(c->a256, nullptr, c->P(nt + "ln.weight"), c->P(nt + "ln.bias"), nullptr, out, R, 256, st);
  });
}

// TorsionAngleHead.forward (layers.py:199-213; linear_3 is registered but unused): out [rows][2], unit norm (clamp 1e-8)
int s2s_torsion_head(s2s_ctx* c, int64_t rows, const float* s_in, float* out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && s_in && out, "s2s_torsion_head: bad argument");
    check_rows(c, rows);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = (int)rows;
    const std::string tp = "translator.torsion_pred.";
    PrecScope ps(c, TC3);
    linear(c, s_in, 256, c->P(tp + "linear_1.weight"), 256, c->P(tp + "linear_1.bias"), c->a256, 256, R, 256, 256, st, 1);
    linear(c, c->a256, 256, c->P(tp + "linear_2.weight"), 256, c->P(tp + "linear_2.bias"), c->b256, 256, R, 256, 256, st, 0, s_in, 256);
    linear(c, c->b256, 256, c->P(tp + "linear_final.weight"), 256, c->P(tp + "linear_final.bias"), c->psi_u, 2, R, 2, 256, st, 0, nullptr, 0, nullptr, nullptr, EXACT);
    psi_finalize(c->psi_u, nullptr, nullptr, out, R, st);
  });
}

// BackboneUpdate.forward of block blk (layers.py:232-241): out [rows][6] = linear(s), exact fp32
int s2s_backbone_update(s2s_ctx* c, int blk, int64_t rows, const float* s_in, float* out, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && s_in && out && blk >= 0 && blk < N_BLK, "s2s_backbone_update: bad argument");
    S2S_CHECK(c->finalized && rows > 0, "s2s_backbone_update: context not finalized");
    const std::string bb = "translator.trunk.bb_update_" + std::to_string(blk) + ".linear.";
    linear(c, s_in, 256, c->P(bb + "weight"), 256, c->P(bb + "bias"), out, 6, (int)rows, 6, 256, (cudaStream_t)stream, 0, nullptr, 0, nullptr, nullptr, EXACT);
  });
}

int s2s_se3_step(int B, int L, const float* rigids_t, const float* rigids_0, const float* residue_mask,
                 const float* diffuse_mask, const float* sched_f, const double* sched_d, const float* rot_noise,
                 const float* trans_noise, float noise_scale, int probability_flow, int mode, double* rot_score,
                 double* trans_score, float* rigids_out, void* stream) {
  return guarded([&] {
    S2S_CHECK(B > 0 && L > 0 && rigids_t && residue_mask && sched_f && sched_d, "s2s_se3_step: bad argument");
    S2S_CHECK(mode == 2 || rigids_0, "s2s_se3_step: rigids_0 required");
    S2S_CHECK(mode == 1 || rigids_out, "s2s_se3_step: rigids_out required");
    S2S_CHECK(mode == 0 || (rot_score && trans_score), "s2s_se3_step: score buffers required");
    S2S_CHECK((rot_score == nullptr) == (trans_score == nullptr), "s2s_se3_step: pass both score buffers or none");
    Se3StepArgs a;
    a.B = B; a.L = L; a.rig_t = rigids_t; a.rig_0 = rigids_0; a.mask = residue_mask; a.diffuse = diffuse_mask;
    a.sched_f = sched_f; a.sched_d = sched_d; a.rot_noise = rot_noise; a.trans_noise = trans_noise;
    a.noise_scale = noise_scale; a.probability_flow = probability_flow; a.mode = mode;
    a.rot_score = rot_score; a.trans_score = trans_score; a.rig_out = rigids_out;
    se3_step(a, (cudaStream_t)stream);
  });
}

int s2s_se3_perturb(int B, int L, const float* rot0, const float* trans0, const float* diffuse_mask, const float* sched_f,
                    const double* cdf, const float* omega_grid, const float* axis_noise, const float* u_noise,
                    const float* trans_noise, float* rigids_out, void* stream) {
  return guarded([&] {
    S2S_CHECK(B > 0 && L > 0 && rot0 && trans0 && sched_f && cdf && omega_grid && axis_noise && u_noise && trans_noise && rigids_out, "s2s_se3_perturb: bad argument");
    Se3PerturbArgs a;
    a.B = B; a.L = L; a.rot0 = rot0; a.trans0 = trans0; a.diffuse = diffuse_mask; a.sched_f = sched_f; a.cdf = cdf;
    a.omega_grid = omega_grid; a.axis_noise = axis_noise; a.u_noise = u_noise; a.trans_noise = trans_noise; a.rig_out = rigids_out;
    se3_perturb(a, (cudaStream_t)stream);
  });
}

int s2s_rng_fill(float* out, int B, int64_t n_per_decoy, uint64_t seed, int64_t first_decoy, int stream_id, int uniform, void* stream) {
  return guarded([&] {
    S2S_CHECK(out && B > 0 && n_per_decoy > 0 && first_decoy >= 0 && stream_id >= 0, "s2s_rng_fill: bad argument");
    philox_fill(out, B, (long)n_per_decoy, seed, first_decoy, (unsigned long long)stream_id, uniform, (cudaStream_t)stream);
  });
}

int s2s_rng_fill_rows(float* out, int B, int64_t n_per_decoy, uint64_t seed, const int64_t* decoy_ids, const int32_t* stream_ids,
                      int uniform, void* stream) {
  return guarded([&] {
    S2S_CHECK(out && B > 0 && n_per_decoy > 0 && decoy_ids && stream_ids, "s2s_rng_fill_rows: bad argument");
    philox_fill(out, B, (long)n_per_decoy, seed, 0, 0, uniform, (cudaStream_t)stream, (const long long*)decoy_ids, (const int*)stream_ids);
  });
}

int s2s_backbone_atoms(s2s_ctx* c, int rows, const float* rigids, const float* psi, const int64_t* aatype, float* atom37,
                       float* atom14, void* stream) {
  return guarded([&] {
    S2S_CHECK(c && rows > 0 && rigids && psi && atom37, "s2s_backbone_atoms: bad argument");
    backbone_atoms(rigids, psi, (const long long*)aatype, c->backbone, atom37, atom14, rows, (cudaStream_t)stream);
  });
}

int s2s_linear_tc(const float* A, const float* W, const float* bias, const float* res, float* C, void* out_hi, void* out_lo,
                  int M, int N, int K, int passes, int relu, void* stream) {
  return guarded([&] {
    S2S_CHECK(A && W && (C || out_hi), "s2s_linear_tc: null argument");
    S2S_CHECK(!out_lo || out_hi, "s2s_linear_tc: out_lo needs out_hi");
    S2S_CHECK(M > 0 && N > 0 && K > 0 && K % 16 == 0 && N % 4 == 0, "s2s_linear_tc: needs K % 16 == 0 and N % 4 == 0");
    S2S_CHECK(passes == 1 || passes == 3, "s2s_linear_tc: passes must be 1 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    bf16 *a_hi, *a_lo, *w_hi, *w_lo;
    S2S_CUDA(cudaMallocAsync(&a_hi, (size_t)M * K * 2, st)); S2S_CUDA(cudaMallocAsync(&a_lo, (size_t)M * K * 2, st));
    S2S_CUDA(cudaMallocAsync(&w_hi, (size_t)N * K * 2, st)); S2S_CUDA(cudaMallocAsync(&w_lo, (size_t)N * K * 2, st));
    split_bf16(A, K, M, K, a_hi, a_lo, st);
    split_bf16(W, K, N, K, w_hi, w_lo, st);
    TcGemm g;
    g.A_hi = a_hi; g.A_lo = a_lo; g.a_rows = M; g.a_cols = K; g.a_pitch = K;
    g.B_hi = w_hi; g.B_lo = w_lo; g.b_rows = N; g.b_cols = K; g.b_pitch = K;
    g.M = M; g.N = N; g.K = K; g.passes = passes; g.relu = relu; g.bias = bias; g.res = res; g.ldres = N;
    g.C = C; g.ldc = N; g.out_hi = (bf16*)out_hi; g.out_lo = (bf16*)out_lo; g.ldo = N;
    gemm_tc(g, st);
    S2S_CUDA(cudaFreeAsync(a_hi, st)); S2S_CUDA(cudaFreeAsync(a_lo, st));
    S2S_CUDA(cudaFreeAsync(w_hi, st)); S2S_CUDA(cudaFreeAsync(w_lo, st));
  });
}

int s2s_linear_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int relu, void* stream) {
  return guarded([&] {
    S2S_CHECK(A && W && C, "s2s_linear_f32: null argument");
    GemmArgs g;
    g.A = A; g.lda = K; g.B = W; g.ldb = K; g.C = C; g.ldc = N; g.bias = bias; g.M = M; g.N = N; g.K = K; g.relu = relu;
    gemm_f32(g, (cudaStream_t)stream);
  });
}

}  // extern "C"
