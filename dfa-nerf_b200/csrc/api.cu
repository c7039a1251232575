// C ABI of libdfn.so (see include/dfn.h): model handle, the MLP entry points and the
// hierarchical render_rays pipeline.  Host code only orchestrates launches on the caller's stream.
#include <stdarg.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "model.h"

namespace dfn {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

// ---- measurement hook: CUDA events around the tcgen05 kernel launches ---------------------------
static constexpr int kMaxProf = 4096;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;
static int g_prof_n = 0;
static double g_prof_macs = 0.0;

bool profile_begin(cudaStream_t st, double macs) {
  if (!g_prof_on || g_prof_n >= kMaxProf) return false;
  if ((int)g_prof_ev.size() < 2 * (g_prof_n + 1)) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
    g_prof_ev.push_back(a);
    g_prof_ev.push_back(b);
  }
  g_prof_macs += macs;
  cudaEventRecord(g_prof_ev[2 * g_prof_n], st);
  return true;
}
void profile_end(cudaStream_t st) {
  cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], st);
  ++g_prof_n;
}

// x[p, :] = [PE(o + d*z) | latent | PE(viewdir)]  (upstream run_network; HELP:42-52, HELP:276)
__global__ void build_inputs_kernel(int64_t P, int S, int64_t ray0_pt, int L, int Lv, int dim_aud,
                                    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                    const float* __restrict__ viewdirs, const float* __restrict__ z_vals,
                                    const float* __restrict__ latent, float* __restrict__ x) {
  const int d_pts = 3 + 6 * L, d_v = 3 + 6 * Lv, D = d_pts + dim_aud + d_v;
  const int64_t n = P * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / D;
    const int j = (int)(i % D);
    const int64_t pt = ray0_pt + p;
    const int64_t ray = pt / S;
    float v;
    if (j < d_pts) {
      const int c = j < 3 ? j : (j - 3) % 3;
      const float xc = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], z_vals[pt]));
      if (j < 3) {
        v = xc;
      } else {
        const int k = (j - 3) / 6, r = (j - 3) % 6;
        const float a = __fmul_rn(xc, pow2i(k));
        v = r < 3 ? sinf(a) : cosf(a);
      }
    } else if (j < d_pts + dim_aud) {
      v = latent[j - d_pts];
    } else {
      const int jj = j - d_pts - dim_aud;
      const int c = jj < 3 ? jj : (jj - 3) % 3;
      const float xc = viewdirs[ray * 3 + c];
      if (jj < 3) {
        v = xc;
      } else {
        const int k = (jj - 3) / 6, r = (jj - 3) % 6;
        const float a = __fmul_rn(xc, pow2i(k));
        v = r < 3 ? sinf(a) : cosf(a);
      }
    }
    x[i] = v;
  }
}

__global__ void z_mid_kernel(int R, int S, const float* __restrict__ z, float* __restrict__ zmid) {
  const int64_t n = (int64_t)R * (S - 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (S - 1)), s = (int)(i % (S - 1));
    zmid[i] = __fmul_rn(0.5f, __fadd_rn(z[(int64_t)r * S + s + 1], z[(int64_t)r * S + s]));
  }
}

static constexpr int64_t kFp32ChunkPoints = 1 << 20;

static int64_t fp32_query_workspace(const dfn_model* m, int64_t n_points) {
  const int64_t c = n_points < kFp32ChunkPoints ? n_points : kFp32ChunkPoints;
  const int in_dim = m->desc.input_ch + m->desc.dim_aud + m->desc.input_ch_views;
  return align256(c * in_dim * 4) + mlp_fp32_workspace_bytes(m, c);
}

static int fp32_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                             const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                             void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const dfn_model_desc& d = m->desc;
  const int64_t n_points = R * S;
  if (workspace_bytes < fp32_query_workspace(m, n_points)) {
    set_error("dfn_query_points(fp32): workspace too small");
    return DFN_E_WORKSPACE;
  }
  const int in_dim = d.input_ch + d.dim_aud + d.input_ch_views;
  const int64_t c = n_points < kFp32ChunkPoints ? n_points : kFp32ChunkPoints;
  float* x = reinterpret_cast<float*>(workspace);
  char* ws2 = reinterpret_cast<char*>(workspace) + align256(c * in_dim * 4);
  for (int64_t p0 = 0; p0 < n_points; p0 += c) {
    const int64_t np = n_points - p0 < c ? n_points - p0 : c;
    int64_t blocks = (np * in_dim + 255) / 256;
    if (blocks > (int64_t)num_sms() * 64) blocks = (int64_t)num_sms() * 64;
    build_inputs_kernel<<<(int)blocks, 256, 0, st>>>(np, S, p0, d.multires, d.multires_views, d.dim_aud, rays_o, rays_d,
                                                    viewdirs, z_vals, latent, x);
    DFN_LAUNCH_CHECK();
    int rc = mlp_fp32_forward(m, np, x, raw + p0 * 4, ws2, workspace_bytes - align256(c * in_dim * 4), st);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace dfn

using namespace dfn;

extern "C" int dfn_abi_version(void) { return DFN_ABI_VERSION; }
extern "C" const char* dfn_last_error(void) { return g_err; }
extern "C" int dfn_last_launch_count(void) { return g_launches; }

extern "C" int dfn_debug_set_impl(int impl) {
  tc_set_impl(impl);
  return 0;
}

extern "C" int dfn_debug_set_pp_flags(int flags) {
  pp_set_flags(flags);
  return 0;
}

extern "C" int dfn_debug_trace(void* dev_buffer, int tiles) {
  tc_set_trace(dev_buffer, tiles);
  return 0;
}

extern "C" int dfn_profile_enable(int on) {
  g_prof_on = on != 0;
  g_prof_n = 0;
  g_prof_macs = 0.0;
  return 0;
}

extern "C" int dfn_profile_collect(double* kernel_ms, int64_t* launches, double* algorithmic_macs) {
  double ms = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    DFN_CUDA(cudaEventSynchronize(g_prof_ev[2 * i + 1]));
    float t = 0.f;
    DFN_CUDA(cudaEventElapsedTime(&t, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]));
    ms += t;
  }
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = g_prof_n;
  if (algorithmic_macs) *algorithmic_macs = g_prof_macs;
  g_prof_n = 0;
  g_prof_macs = 0.0;
  return 0;
}

extern "C" int dfn_model_create(const dfn_model_desc* desc, dfn_model** out) {
  DFN_CHECK_ARG(desc && out, "dfn_model_create: null argument");
  DFN_CHECK_ARG(desc->kind == DFN_MODEL_FACENERF || desc->kind == DFN_MODEL_NERF, "dfn_model_create: unknown kind %d", desc->kind);
  DFN_CHECK_ARG(desc->D >= 2 && desc->D <= 16 && desc->W >= 16 && desc->W % 2 == 0, "dfn_model_create: bad D/W");
  DFN_CHECK_ARG(desc->input_ch == 3 + 6 * desc->multires && desc->input_ch_views == 3 + 6 * desc->multires_views,
                "dfn_model_create: input_ch must equal 3+6*multires");
  DFN_CHECK_ARG(desc->kind == DFN_MODEL_FACENERF ? desc->dim_aud > 0 : desc->dim_aud == 0, "dfn_model_create: dim_aud");
  dfn_model* m = new (std::nothrow) dfn_model();
  if (!m) {
    set_error("dfn_model_create: out of memory");
    return DFN_E_STATE;
  }
  m->desc = *desc;
  m->n_views = desc->kind == DFN_MODEL_FACENERF ? 1 + desc->D / 4 : 1;
  *out = m;
  return 0;
}

extern "C" void dfn_model_destroy(dfn_model* m) {
  if (!m) return;
  cudaFree(m->fp32_blob);
  tc_free_model(m);
  delete m;
}

extern "C" int dfn_model_num_tensors(const dfn_model* m) { return m ? 2 * (m->desc.D + m->n_views + 3) : 0; }

extern "C" int dfn_model_load(dfn_model* m, const float* const* t, int n_tensors, void* stream) {
  DFN_CHECK_ARG(m && t, "dfn_model_load: null argument");
  DFN_CHECK_ARG(n_tensors == dfn_model_num_tensors(m), "dfn_model_load: expected %d tensors, got %d",
                dfn_model_num_tensors(m), n_tensors);
  for (int i = 0; i < n_tensors; ++i) DFN_CHECK_ARG(t[i] != nullptr, "dfn_model_load: tensor %d is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  const dfn_model_desc& d = m->desc;
  const int n_pts = d.input_ch + d.dim_aud, W = d.W, Wh = W / 2;
  // (in, out) of every Linear in load order
  std::vector<std::pair<int, int>> shp;
  for (int i = 0; i < d.D; ++i) shp.push_back({i == 0 ? n_pts : (i - 1 == d.skip ? W + n_pts : W), W});
  for (int i = 0; i < m->n_views; ++i) shp.push_back({i == 0 ? W + d.input_ch_views : Wh, Wh});
  shp.push_back({W, W});   // feature_linear
  shp.push_back({W, 1});   // alpha_linear
  shp.push_back({Wh, 3});  // rgb_linear
  size_t total = 0;
  for (auto& s : shp) total += (size_t)s.first * s.second + s.second;
  std::vector<float> host(total);
  std::vector<size_t> offs;
  size_t o = 0;
  for (size_t i = 0; i < shp.size(); ++i) {
    offs.push_back(o);
    memcpy(&host[o], t[2 * i], (size_t)shp[i].first * shp[i].second * 4);
    o += (size_t)shp[i].first * shp[i].second;
    memcpy(&host[o], t[2 * i + 1], (size_t)shp[i].second * 4);
    o += shp[i].second;
  }
  cudaFree(m->fp32_blob);
  m->fp32_blob = nullptr;
  tc_free_model(m);
  m->loaded = false;
  DFN_CUDA(cudaMalloc(&m->fp32_blob, total * 4));
  DFN_CUDA(cudaMemcpyAsync(m->fp32_blob, host.data(), total * 4, cudaMemcpyHostToDevice, st));
  DFN_CUDA(cudaStreamSynchronize(st));
  auto mk = [&](size_t i) {
    Fp32Layer L;
    L.w = m->fp32_blob + offs[i];
    L.b = L.w + (size_t)shp[i].first * shp[i].second;
    L.in = shp[i].first;
    L.out = shp[i].second;
    return L;
  };
  for (int i = 0; i < d.D; ++i) m->pts[i] = mk(i);
  for (int i = 0; i < m->n_views; ++i) m->views[i] = mk(d.D + i);
  m->feature = mk(d.D + m->n_views);
  m->alpha = mk(d.D + m->n_views + 1);
  m->rgb = mk(d.D + m->n_views + 2);
  // tcgen05 layouts (optional: shapes outside its coverage keep the fp32 path only)
  int rc = tc_pack_model(m, t, st);
  if (rc != 0 && rc != DFN_E_UNSUPPORTED) return rc;
  if (rc == DFN_E_UNSUPPORTED) tc_free_model(m);
  m->loaded = true;
  return 0;
}

// Host-only: the layer program dfn_model_load builds for the tcgen05 kernels, as dense fp32 (no CUDA calls).
extern "C" int dfn_model_program_host(const dfn_model* m, const float* const* t, int n_tensors, int max_layers,
                                      dfn_layer_info* layers, int* n_layers, float* weights, float* bias, int* fold_layer,
                                      float* fold_w, float* view_w, float* view_b) {
  DFN_CHECK_ARG(m && t && layers && n_layers && weights && bias && fold_layer && fold_w && view_w && view_b,
                "dfn_model_program_host: null argument");
  DFN_CHECK_ARG(n_tensors == dfn_model_num_tensors(m) && max_layers >= TC_MAX_LAYERS,
                "dfn_model_program_host: expected %d tensors and max_layers >= %d", dfn_model_num_tensors(m), TC_MAX_LAYERS);
  dfn_model tmp;
  tmp.desc = m->desc;
  tmp.n_views = m->n_views;
  TcHostDump dump;
  int rc = tc_pack_model(&tmp, t, nullptr, &dump);
  if (rc) return rc;
  *n_layers = tmp.prog.n_layers;
  for (int l = 0; l < tmp.prog.n_layers; ++l) {
    const TcLayer& L = tmp.prog.layers[l];
    layers[l].n = L.n;
    layers[l].nkb = L.nkb;
    layers[l].epi = L.epi;
    layers[l].flags = L.flags;
    for (int k = 0; k < 6; ++k) layers[l].kb[k] = L.kb[k];
  }
  memset(weights, 0, (size_t)max_layers * TC_BIAS_STRIDE * 6 * 64 * 4);
  memcpy(weights, dump.dense.data(), dump.dense.size() * 4);
  memcpy(bias, dump.bias.data(), dump.bias.size() * 4);
  fold_layer[0] = tmp.prog.fold_layer[0];
  fold_layer[1] = tmp.prog.fold_layer[1];
  memcpy(fold_w, dump.fold_w.data(), dump.fold_w.size() * 4);
  memcpy(view_w, dump.view_w.data(), dump.view_w.size() * 4);
  memcpy(view_b, dump.view_b.data(), dump.view_b.size() * 4);
  return 0;
}

extern "C" int64_t dfn_mlp_workspace_bytes(const dfn_model* m, int64_t P) {
  return m ? mlp_fp32_workspace_bytes(m, P) : 0;
}

extern "C" int dfn_mlp_forward(const dfn_model* m, int64_t P, const float* x, float* out, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  DFN_CHECK_ARG(m && x && out && workspace && P > 0, "dfn_mlp_forward: bad argument");
  if (!m->loaded) {
    set_error("dfn_mlp_forward: model has no weights");
    return DFN_E_STATE;
  }
  reset_launch_count();
  return mlp_fp32_forward(m, P, x, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int64_t dfn_query_workspace_bytes(const dfn_model* m, int64_t R, int S, int precision) {
  if (!m) return 0;
  if (precision == DFN_PREC_FP32) return fp32_query_workspace(m, R * S);
  return tc_query_workspace_bytes(m, R, S);
}

static int query_points_impl(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                             const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                             int precision, void* workspace, int64_t workspace_bytes, cudaStream_t st, const void* pre = nullptr) {
  DFN_CHECK_ARG(m && R > 0 && S > 0 && rays_o && rays_d && viewdirs && z_vals && raw && workspace,
                "dfn_query_points: bad argument");
  if (!m->loaded) {
    set_error("dfn_query_points: model has no weights");
    return DFN_E_STATE;
  }
  DFN_CHECK_ARG(m->desc.dim_aud == 0 || latent != nullptr, "dfn_query_points: this model needs a latent of %d floats",
                m->desc.dim_aud);
  if (precision == DFN_PREC_FP32)
    return fp32_query_points(m, R, S, rays_o, rays_d, viewdirs, z_vals, latent, raw, workspace, workspace_bytes, st);
  if (precision != DFN_PREC_BF16 && precision != DFN_PREC_BF16X3 && precision != DFN_PREC_FP16 && precision != DFN_PREC_FP16X3M) {
    set_error("dfn_query_points: unknown precision %d", precision);
    return DFN_E_ARG;
  }
  if (m->tc_hi == nullptr) {
    set_error("dfn_query_points: this model shape has no tcgen05 path; use DFN_PREC_FP32");
    return DFN_E_UNSUPPORTED;
  }
  return tc_query_points(m, R, S, rays_o, rays_d, viewdirs, z_vals, latent, raw, precision, workspace, workspace_bytes, st, pre);
}

extern "C" int dfn_query_points(const dfn_model* m, int64_t R, int S, const float* rays_o, const float* rays_d,
                                const float* viewdirs, const float* z_vals, const float* latent, float* raw,
                                int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  reset_launch_count();
  return query_points_impl(m, R, S, rays_o, rays_d, viewdirs, z_vals, latent, raw, precision, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------- render_rays
struct RenderWs {
  int64_t z0, raw0, w0, zmid, zs, zall, raw1, query, pre_a, pre_b, total;
};

// query scratch and prepared-bias regions are sized for the LARGER of the two networks (a fine network of another width or latent size
// runs in the same regions)
static RenderWs render_layout(const dfn_model* m, const dfn_model* m2, int64_t R, int Nc, int Nf, int precision) {
  RenderWs w;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    int64_t at = o;
    o += align256(bytes);
    return at;
  };
  const int Nt = Nc + Nf;
  w.z0 = take(R * Nc * 4);
  w.raw0 = take(R * Nc * 16);
  w.w0 = take(R * Nc * 4);
  w.zmid = take(R * (Nc > 1 ? Nc - 1 : 1) * 4);
  w.zs = take(R * (Nf > 0 ? Nf : 1) * 4);
  w.zall = take(R * Nt * 4);
  w.raw1 = take(Nf > 0 ? R * Nt * 16 : 16);
  int64_t q = dfn_query_workspace_bytes(m, R, Nc, precision);
  int64_t pb = precision == DFN_PREC_FP32 ? 16 : tc_prep_bytes(m, R);   // both networks' prepared biases (tc_prep_launch)
  const dfn_model* mf = m2 ? m2 : m;
  if (Nf > 0) {
    const int64_t qf = dfn_query_workspace_bytes(mf, R, Nt, precision);
    if (qf > q) q = qf;
    if (precision != DFN_PREC_FP32 && tc_prep_bytes(mf, R) > pb) pb = tc_prep_bytes(mf, R);
  }
  w.query = take(q);
  w.pre_a = take(pb);
  w.pre_b = take(pb);
  w.total = o;
  return w;
}

extern "C" int64_t dfn_render_workspace_bytes(const dfn_model* coarse, int64_t R, int N_samples, int N_importance,
                                              int precision) {
  if (!coarse || R <= 0 || N_samples <= 0 || N_importance < 0) return 0;
  return render_layout(coarse, nullptr, R, N_samples, N_importance, precision).total;
}

extern "C" int64_t dfn_render_workspace_bytes2(const dfn_model* coarse, const dfn_model* fine, int64_t R, int N_samples,
                                               int N_importance, int precision) {
  if (!coarse || R <= 0 || N_samples <= 0 || N_importance < 0) return 0;
  return render_layout(coarse, fine, R, N_samples, N_importance, precision).total;
}

extern "C" int dfn_render_rays(const dfn_model* coarse, const dfn_model* fine, int64_t R, int N_samples,
                               int N_importance, const dfn_render_io* io, int white_bkgd, int precision,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  reset_launch_count();
  DFN_CHECK_ARG(coarse && io && workspace && R > 0 && R < (1ll << 31) && N_samples >= 2 && N_importance >= 0,
                "dfn_render_rays: bad argument");
  DFN_CHECK_ARG(io->rays_o && io->rays_d && io->viewdirs && io->near && io->far && io->t_vals,
                "dfn_render_rays: rays_o/rays_d/viewdirs/near/far/t_vals are required");
  DFN_CHECK_ARG(N_importance == 0 || io->u_vals || io->z_samples_in, "dfn_render_rays: u_vals required for the fine pass");
  DFN_CHECK_ARG(N_samples + N_importance <= 256, "dfn_render_rays: at most 256 samples per ray");
  if (fine == nullptr) fine = coarse;
  cudaStream_t st = (cudaStream_t)stream;
  const int Nc = N_samples, Nf = N_importance, Nt = Nc + Nf;
  const RenderWs L = render_layout(coarse, fine, R, Nc, Nf, precision);
  if (workspace_bytes < L.total) {
    set_error("dfn_render_rays: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)L.total);
    return DFN_E_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  float* z0 = reinterpret_cast<float*>(ws + L.z0);
  float* raw0 = reinterpret_cast<float*>(ws + L.raw0);
  float* w0 = reinterpret_cast<float*>(ws + L.w0);
  float* zmid = reinterpret_cast<float*>(ws + L.zmid);
  float* zs = reinterpret_cast<float*>(ws + L.zs);
  float* zall = io->z_vals_out && Nf > 0 ? io->z_vals_out : reinterpret_cast<float*>(ws + L.zall);
  float* raw1 = reinterpret_cast<float*>(ws + L.raw1);
  void* qws = ws + L.query;
  const int64_t qbytes = L.pre_a - L.query;
  const int Ri = (int)R;
  int rc;
#define STEP(call)        \
  do {                    \
    rc = (call);          \
    if (rc) return rc;    \
  } while (0)

  // one preparation launch for the frame: both networks' folded biases and per-ray view-bias rows + the coarse depths
  const void* pre_c = nullptr;
  const void* pre_f = nullptr;
  rc = DFN_E_UNSUPPORTED;
  if (precision != DFN_PREC_FP32 && Nf > 0 && coarse->loaded && fine->loaded)
    rc = tc_prep_launch(coarse, fine, R, Nc, io->viewdirs, io->latent, io->t_vals, io->near, io->far, io->perturb_rand, z0, ws + L.pre_a,
                        ws + L.pre_b, st);
  if (rc == 0) {
    pre_c = ws + L.pre_a;
    pre_f = ws + L.pre_b;
  } else if (rc != DFN_E_UNSUPPORTED) {
    return rc;
  } else {
    STEP(dfn_z_vals(Ri, Nc, io->t_vals, io->near, io->far, io->perturb_rand, z0, st));
  }
  STEP(query_points_impl(coarse, R, Nc, io->rays_o, io->rays_d, io->viewdirs, z0, io->latent, raw0, precision, qws,
                         qbytes, st, pre_c));
  if (Nf == 0) {
    STEP(launch_raw2outputs(Ri, Nc, raw0, z0, io->rays_d, io->bc_rgb, 0, white_bkgd, 1e10f, io->rgb_map, io->disp_map,
                            io->acc_map, nullptr, nullptr, io->last_weight, st));
    if (io->z_vals_out) DFN_CUDA(cudaMemcpyAsync(io->z_vals_out, z0, (size_t)R * Nc * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    if (Nc >= 3 && Nc <= 128) {
      // raw2outputs(coarse) -> z_mid -> sample_pdf -> sort-merge in ONE launch; weights, midpoints, cdf and the new samples stay
      // in shared memory (stages.cu: coarse_to_fine_kernel, bit-identical to the chain below)
      STEP(dfn_coarse_to_fine(Ri, Nc, Nf, raw0, z0, io->rays_d, io->bc_rgb, white_bkgd, 1e10f, io->u_vals, io->u_per_ray ? 1 : 0,
                              io->z_samples_in, io->rgb0, io->z_samples_out, zall, st));
    } else {
      STEP(launch_raw2outputs(Ri, Nc, raw0, z0, io->rays_d, io->bc_rgb, 0, white_bkgd, 1e10f, io->rgb0, nullptr, nullptr, w0,
                              nullptr, nullptr, st));
      const float* zsamp = io->z_samples_in;
      if (zsamp == nullptr) {
        int64_t blocks = ((int64_t)R * (Nc - 1) + 255) / 256;
        if (blocks > (int64_t)num_sms() * 32) blocks = (int64_t)num_sms() * 32;
        z_mid_kernel<<<(int)blocks, 256, 0, st>>>(Ri, Nc, z0, zmid);
        DFN_LAUNCH_CHECK();
        float* zs_out = io->z_samples_out ? io->z_samples_out : zs;
        STEP(dfn_sample_pdf(Ri, Nc - 1, zmid, w0 + 1, Nc, Nf, io->u_vals, io->u_per_ray ? 1 : 0, zs_out, nullptr, st));
        zsamp = zs_out;
      } else if (io->z_samples_out && io->z_samples_out != zsamp) {
        DFN_CUDA(cudaMemcpyAsync(io->z_samples_out, zsamp, (size_t)R * Nf * 4, cudaMemcpyDeviceToDevice, st));
      }
      STEP(dfn_sort_merge(Ri, Nc, z0, Nf, zsamp, zall, st));
    }
    STEP(query_points_impl(fine, R, Nt, io->rays_o, io->rays_d, io->viewdirs, zall, io->latent, raw1, precision, qws,
                           qbytes, st, pre_f));
    STEP(launch_raw2outputs(Ri, Nt, raw1, zall, io->rays_d, io->bc_rgb, 0, white_bkgd, 1e10f, io->rgb_map, io->disp_map,
                            io->acc_map, nullptr, nullptr, io->last_weight, st));
  }
#undef STEP
  return 0;
}
