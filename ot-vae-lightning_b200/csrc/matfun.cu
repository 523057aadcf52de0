// K3-K6: SPD matrix square roots by coupled Newton-Schulz, Gaussian W2^2 and the transport operator.
// Reference: ot/matrix_utils.py:37-76 ; ot/w2_utils.py:40-80,756-768.
//
// All d x d products run through gemm_any: fp32 work -> tcgen05 3xTF32 (or FFMA for odd shapes), fp64 work -> DFMA.
// Products are true NN products (see gemm.cuh: substituting B^T for the nearly-symmetric B destabilises the iteration).
//
// Precision policy: the fp32-accurate engine is tried first.  Its accuracy is ~1e-7 * cond(A), so when the iteration
// needs more than NS_F32_MAX_ITERS steps (which happens when lambda_min/||A||_F is below ~1e-4) and the caller's
// data is fp64 (the reference default), the whole computation is repeated in fp64 on the same device.
#include <mutex>
#include <vector>
#include <cstdlib>
#include "gemm.cuh"
#include "otk_ptx.cuh"

namespace otk {

constexpr int NS_MAX_ITERS = 96;       // size of the residual log
constexpr int NS_F32_MAX_ITERS = 32;   // fp32 engine: beyond this the input is too ill-conditioned for fp32
constexpr int NS_F64_MAX_ITERS = 90;
// iterations-to-converge grows like log_2.25(||A||_F / lambda_min): used as the conditioning estimate that decides
// when the fp32 engine's ~1e-7 * cond error would exceed the parity tolerance (measured: T err 3e-5 at cond 1e2,
// 3e-3 at cond 1e4)
constexpr int NS_F32_OPERATOR_ITERS = 14;   // (counts include the NS_ACCEL_STEPS accelerated iterations, each worth ~3 classical ones)
constexpr double NS_F32_RICCATI_TOL = 2e-4;   // ||T Cs T - Ct||_F / ||Ct||_F accepted from the fp32 engine
// (calibration, cfg5 sweep on a B200: the relative error of the fp32 map against the fp64 engine is 3 - 4.5 x its Riccati
// residual - 1.2e-4 / 4.6e-5 at d = 1024, 4.1e-4 / 1.3e-4 at 2048, 1.6e-3 / 3.4e-4 at 4096 with un-split K - so 2e-4 keeps
// the map inside the 1e-3 budget for every width)
static inline double riccati_tol(int64_t) { return NS_F32_RICCATI_TOL; }
constexpr int64_t NS_SMALL_DIM = 48;          // fp64 data with dim <= 48: the DFMA engine is as fast and exact (d = 64: 1.46 ms there, 0.45 ms on the tcgen05 graph)
constexpr int NS_F32_IROOT_ITERS = 12;
// a root alone loses ~1e-7 * sqrt(cond): accepted from the fp32 engine up to this many iterations (lambda_min / c down to
// ~1e-6); beyond that - ill-conditioned or rank-deficient input, whose noise-level eigenvalues the fp32 engine would
// "converge" on - the fp64 engine takes over
constexpr int NS_F32_ROOT_ITERS = 16;
// Accelerated start.  The classical step multiplies an eigenvalue p << 1 of Z Y by 2.25 per iteration (3 products).  The
// first NS_ACCEL_STEPS iterations use the degree-2 polynomial in M = Z Y instead,
//     T = a I + b M + c M^2 ,  Y <- Y T ,  Z <- T Z     (p <- p T(p)^2 : a factor a^2 = 11.9 per iteration, 4 products),
// with the quintic coefficients published for Newton-Schulz orthogonalisation (Jordan et al., "Muon", 2024: x <- a x +
// b x^3 + c x^5 on the singular values, here with x^2 = p).  It maps (0, 1.4] into [0.5, 1.25] but does not converge to 1,
// so the classical (quadratically convergent, self-correcting) steps finish the job.  Measured on the fp32 simulation
// (scratch/ns_accel.py): d = 512, cond 1e2: 27 products instead of 39, cond 1e4: 39 instead of 54, same or better accuracy.
constexpr int NS_ACCEL_STEPS = 3;
constexpr int NS_ACCEL_STEPS_ESCALATED = 8;   // fp64 engine re-running what the fp32 engine found ill-conditioned / singular
constexpr double NS_ACC_A = 3.4445, NS_ACC_B = -4.7750, NS_ACC_C = 2.0315;
static thread_local int g_ns_accel_steps = NS_ACCEL_STEPS;
// The fp64 engine adds NS_F64_REL_RIDGE * ||A||_F to the diagonal: ten fp64 ulps of the norm, below the iteration's own
// round-off for any PD input, but it turns an exactly singular SPSD matrix (a rank-deficient covariance, the 'spsd'
// arguments of the reference: w2_utils.py:73-76, 423-426; FID with fewer samples than features) into one the iteration
// converges on in ~48 steps with a root error of sqrt(1e-15) - the reference's eigh-based sqrtm is finite there too.
constexpr double NS_F64_REL_RIDGE = 1e-15;

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < blockDim.x / 32 ? red[threadIdx.x] : 0.0;
    r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// Scale of the iteration: c[l] >= lambda_max(A_l + ridge I), the smaller of the Frobenius norm and the maximum absolute row
// sum (both bound the spectral radius; for a covariance with a spread spectrum the row-sum norm is several times tighter
// than the Frobenius norm, and every factor 2.25 saved is one Newton-Schulz iteration).  One warp per row: sum of squares
// with an fp64 atomic into c[l], row sum with an atomic max (bit pattern of a non-negative double) into cmax[l].
__global__ void norm_rows_kernel(const void* a, int dt, int64_t dim, double ridge, double* c, double* cmax) {
  const int64_t l = blockIdx.y;
  const int lane = threadIdx.x % 32;
  double sq_tot = 0, mx = 0;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; row < dim; row += (int64_t)gridDim.x * (blockDim.x / 32)) {
    double sq = 0, ab = 0;
    for (int64_t j = lane; j < dim; j += 32) {
      double v = load_real(a, l * dim * dim + row * dim + j, dt);
      if (j == row) v += ridge;
      sq += v * v;
      ab += fabs(v);
    }
    sq = warp_sum(sq);
    ab = warp_sum(ab);
    sq_tot += sq;
    mx = ab > mx ? ab : mx;
  }
  if (lane == 0) {
    atomicAdd(&c[l], sq_tot);
    atomicMax(reinterpret_cast<unsigned long long*>(&cmax[l]), (unsigned long long)__double_as_longlong(mx));
  }
}
// On entry c[l] = ||A + ridge I||_F^2 and ridge_l[l] = its maximum absolute row sum.  c0 = min(sqrt(c[l]), row-sum norm);
// ridge_l[l] <- ridge + rel * c0 ;  c[l] <- c0 * (1 + rel)  (still an upper bound of lambda_max(A + ridge_l I))
__global__ void ns_ridge_kernel(double* c, int64_t L, double ridge, double rel, double* ridge_l) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < L) {
    const double fro = sqrt(c[e]), rs = ridge_l[e];
    const double c0 = (rs > 0 && rs < fro) ? rs : fro;
    ridge_l[e] = ridge + rel * c0;
    c[e] = c0 > 0 ? c0 * (1.0 + rel) : 1.0;   // the zero matrix: Y stays 0, the solve reports not-converged
  }
}
static int frob_norm(const void* a, int dt, int64_t L, int64_t dim, double ridge, double rel, double* c, double* ridge_l,
                     cudaStream_t st) {
  OTK_CUDA(cudaMemsetAsync(c, 0, (size_t)L * 8, st));
  OTK_CUDA(cudaMemsetAsync(ridge_l, 0, (size_t)L * 8, st));
  int64_t bx = ceil_div(dim, 8);
  if (bx > 128) bx = 128;
  if (bx < 1) bx = 1;
  norm_rows_kernel<<<dim3((unsigned)bx, (unsigned)L), 256, 0, st>>>(a, dt, dim, ridge, c, ridge_l);
  ns_ridge_kernel<<<(unsigned)ceil_div(L, 256), 256, 0, st>>>(c, L, ridge, rel, ridge_l);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// Y = (A + ridge I)/c, Z = I
template <typename W>
__global__ void ns_init_kernel(const void* a, int dt, int64_t L, int64_t dim, const double* ridge_l, const double* c, W* Y, W* Z) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim);
    bool diag = (r / dim == r % dim);
    double v = load_real(a, e, dt) + (diag ? ridge_l[l] : 0.0);
    Y[e] = (W)(v / c[l]);
    Z[e] = diag ? W(1) : W(0);
  }
}

// out = s0 * c[l]^pw * (in + in^T)/2 + s1 * c[l]^pw1 * (in2 + in2^T)/2 + diag_add * I   (in2 optional), cast to out_dt
template <typename W>
__global__ void sym_scale_kernel(const W* in, const W* in2, int64_t L, int64_t dim, const double* c, double pw, double s0,
                                 double pw1, double s1, double diag_add, void* out, int out_dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    int64_t et = l * dim * dim + j * dim + i;
    double v = 0.5 * ((double)in[e] + (double)in[et]) * s0 * (c ? pow(c[l], pw) : 1.0);
    if (in2) v += 0.5 * ((double)in2[e] + (double)in2[et]) * s1 * (c ? pow(c[l], pw1) : 1.0);
    if (i == j) v += diag_add;
    store_real(out, e, out_dt, v);
  }
}

template <typename W>
__global__ void cast_kernel(const void* a, int dt, int64_t n, W* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = (W)load_real(a, e, dt);
}

// w2[l] = |ms - mt|^2 + tr(Cs) + tr(Ct) - 2 * sqrt(c[l]) * tr(Y_l)      (one block per l)
template <typename W>
__global__ void w2_trace_kernel(const void* ms, const void* mt, const void* cs, const void* ct, int dt, int64_t dim,
                                const W* Y, const double* c, double* w2) {
  __shared__ double red[32];
  const int64_t l = blockIdx.x;
  double acc = 0;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    double dm = load_real(ms, l * dim + i, dt) - load_real(mt, l * dim + i, dt);
    int64_t dgl = l * dim * dim + i * dim + i;
    acc += dm * dm + load_real(cs, dgl, dt) + load_real(ct, dgl, dt) - 2.0 * sqrt(c[l]) * (double)Y[dgl];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) w2[l] = acc;
}

// acc[2l] += || R_l - target_l ||_F^2 , acc[2l+1] += || target_l ||_F^2    (grid = blocks x L; acc zeroed by the caller)
__global__ void rel_residual_kernel(const float* R, const float* target, int64_t dim, double* acc) {
  __shared__ double red[32];
  const int64_t l = blockIdx.y;
  double num = 0, den = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < dim * dim; e += (int64_t)gridDim.x * blockDim.x) {
    double t = target[l * dim * dim + e], df = (double)R[l * dim * dim + e] - t;
    num += df * df;
    den += t * t;
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) { atomicAdd(&acc[2 * l], num); atomicAdd(&acc[2 * l + 1], den); }
}

// Device-side control of the Newton-Schulz loop.  ctrl[0] = number of iterations to execute (kernels of iteration
// k >= ctrl[0] retire immediately), ctrl[1] = verdict, ctrl[2] = diverged.  Runs after the Z*Y product of iteration k,
// whose epilogue accumulated resid_k[l] = ||I - Z_k Y_k||_F^2; mirrors the stopping rule the host used to apply after a
// synchronisation per iteration.
__global__ void ns_ctrl_init_kernel(int* ctrl, int max_iters, int verdict0) {
  if (threadIdx.x == 0) { ctrl[0] = max_iters; ctrl[1] = verdict0; ctrl[2] = 0; ctrl[3] = 0; }
}
__global__ void ns_ctrl_kernel(const double* resid_k, int64_t L, int k, int max_iters, double tol_done, double tol_near,
                               int* ctrl) {
  __shared__ double red[32];
  if (k >= ctrl[0]) return;
  if (!(k >= 5 || k + 1 == max_iters)) return;
  double worst = 0;
  for (int64_t l = threadIdx.x; l < L; l += blockDim.x) {
    const double r = resid_k[l];
    const double v = (r == r && r <= 1e30) ? r : 1e300;
    worst = v > worst ? v : worst;
  }
  worst = warp_max(worst);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)blockDim.x / 32; ++w) worst = red[w] > worst ? red[w] : worst;
    if (worst >= 1e300) { ctrl[0] = k + 1; ctrl[1] = 1; ctrl[2] = 1; }          // diverged: numerically indefinite input
    else if (worst < tol_done) { if (k + 1 < ctrl[0]) ctrl[0] = k + 1; ctrl[1] = 0; }
    else if (worst < tol_near) { if (k + 2 < ctrl[0]) ctrl[0] = k + 2; ctrl[1] = 0; }
  }
}

template <typename T>
static inline GemmArgs<T> with_scratch(GemmArgs<T> g, T* scratch) { g.scratch = scratch; return g; }

static inline unsigned ew_grid(int64_t total) {
  int64_t b = ceil_div(total, 256), cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

template <typename W>
struct NsWork {
  W *Y[2], *Z[2], *T, *M;
  // TF32 hi/lo planes of the iterates (fp32 / tcgen05 engine only): no conversion pass between chained GEMMs
  W *Yh[2], *Yl[2], *Zh[2], *Zl[2], *Th, *Tl, *Mh, *Ml;
  W* scratch;         // 4 planes: operand splits for the products outside the iteration
  double *c, *resid;  // c [L]; resid [NS_MAX_ITERS][L]
  double* ridge_l;    // [L] diagonal shift of the current solve
  int* ctrl;          // device-side loop control (ns_ctrl_kernel)
  static constexpr int kPlanes = sizeof(W) == 4 ? 6 + 12 + 4 : 6;
  static size_t bytes(int64_t L, int64_t d) {
    return kPlanes * align_up((size_t)L * d * d * sizeof(W), 256) + 2 * align_up((size_t)L * 8, 256) +
           align_up((size_t)NS_MAX_ITERS * L * 8, 256) + 256;
  }
  void carve(Arena& ar, int64_t L, int64_t d) {
    const size_t n = (size_t)L * d * d;
    for (int i = 0; i < 2; ++i) { Y[i] = ar.take<W>(n); Z[i] = ar.take<W>(n); }
    T = ar.take<W>(n);
    M = ar.take<W>(n);
    if (sizeof(W) == 4) {
      for (int i = 0; i < 2; ++i) { Yh[i] = ar.take<W>(n); Yl[i] = ar.take<W>(n); Zh[i] = ar.take<W>(n); Zl[i] = ar.take<W>(n); }
      Th = ar.take<W>(n); Tl = ar.take<W>(n);
      Mh = ar.take<W>(n); Ml = ar.take<W>(n);
      scratch = ar.take<W>(4 * n);
    } else {
      scratch = nullptr;
    }
    c = ar.take<double>((size_t)L);
    ridge_l = ar.take<double>((size_t)L);
    resid = ar.take<double>((size_t)NS_MAX_ITERS * L);
    ctrl = ar.take<int>(16);
  }
};
static size_t ns_work_bytes(int64_t L, int64_t d) {
  size_t a = NsWork<float>::bytes(L, d), b = NsWork<double>::bytes(L, d);
  return a > b ? a : b;
}

// plane-mode helpers (fp32 engine on tcgen05)
__global__ void ns_init_planes_kernel(const void* a, int dt, int64_t L, int64_t dim, const double* ridge_l, const double* c, float* Yh,
                                      float* Yl, float* Zh, float* Zl) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim);
    bool diag = (r / dim == r % dim);
    float v = (float)((load_real(a, e, dt) + (diag ? ridge_l[l] : 0.0)) / c[l]);
    float h, lo;
    ptx::split_tf32(v, h, lo);
    Yh[e] = h; Yl[e] = lo;
    Zh[e] = diag ? 1.f : 0.f; Zl[e] = 0.f;
  }
}
__global__ void join_planes_kernel(const float* hi, const float* lo, int64_t n, float* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = hi[e] + lo[e];
}
static bool ns_planes_eligible(int64_t d) { return d >= 64 && d % 4 == 0; }
static GemmArgs<float> plane_args(const float* Ah, const float* Al, const float* Bh, const float* Bl, float* Ch, float* Cl,
                                  int64_t d, float alpha, float diag, double* resid) {
  GemmArgs<float> g = nn_args_t<float>(Ah, Bh, nullptr, d, d * d, alpha);
  g.A_lo = Al; g.B_lo = Bl; g.C_hi = Ch; g.C_lo = Cl; g.diag_add = diag; g.resid = resid;
  return g;
}
static int plane_gemm2(const GemmArgs<float>& g0, const GemmArgs<float>* g1, int64_t L, const int* ctrl, int k, cudaStream_t st,
                       const NsCtrlEval* eval = nullptr) {
  int r = gemm_umma_dual(g0, g1, L, ctrl, k, st, eval);
  if (r == 0) { set_last_error_msg("sqrtm: tcgen05 engine rejected an eligible shape"); return OTK_ERR_CUDA; }
  return r < 0 ? r : OTK_OK;
}

enum { NS_CONVERGED = 0, NS_SLOW = 1 };

// ---------------------------------------------------------------------------------------------------------------------
// CUDA-graph cache for batches of Newton-Schulz iterations on the tcgen05 path.  A batch is ~3 launches per iteration,
// each with ~1 KB of __grid_constant__ tensor maps encoded on the host: at d <= 512 the chain is bound by the CPU's launch
// rate, not by the GPU.  All launches of a batch depend only on (workspace pointers, L, d, first iteration, count,
// budget) and carry their stopping rule on the device (ctrl), so a batch that is seen a second time with the same key is
// captured once (on a private stream: torch's legacy default stream cannot be captured) and replayed afterwards with one
// cudaGraphLaunch on the caller's stream.
// ---------------------------------------------------------------------------------------------------------------------
unsigned long long launches();
struct NsGraphKey {
  const void *ctrl, *planes;
  int64_t L, d;
  int first, n, max_iters, adaptive, dev;
  bool operator==(const NsGraphKey& o) const {
    return ctrl == o.ctrl && planes == o.planes && L == o.L && d == o.d && first == o.first && n == o.n &&
           max_iters == o.max_iters && adaptive == o.adaptive && dev == o.dev;
  }
};
struct NsGraphEntry { NsGraphKey key; cudaGraphExec_t exec; int launches; int state; };   // state 0 seen once, 1 ready, -1 unusable
static std::mutex g_ns_graph_mu;
static std::vector<NsGraphEntry> g_ns_graphs;
static cudaStream_t g_ns_capture_stream[64] = {nullptr};
constexpr size_t NS_GRAPH_CACHE = 32;

template <typename Enqueue>
static int ns_batch(const NsGraphKey& key, cudaStream_t st, Enqueue&& enqueue) {
  static const bool graphs_on = [] { const char* e = getenv("OTK_NS_GRAPHS"); return !(e && e[0] == '0'); }();   // tuning aid
  if (!graphs_on || key.dev < 0 || key.dev >= 64) return enqueue(st);
  std::lock_guard<std::mutex> lock(g_ns_graph_mu);
  NsGraphEntry* hit = nullptr;
  for (auto& e : g_ns_graphs) if (e.key == key) { hit = &e; break; }
  if (!hit) {                                                   // first sighting: plain launches, remember the key
    if (g_ns_graphs.size() >= NS_GRAPH_CACHE) {
      if (g_ns_graphs.front().exec) cudaGraphExecDestroy(g_ns_graphs.front().exec);
      g_ns_graphs.erase(g_ns_graphs.begin());
    }
    g_ns_graphs.push_back(NsGraphEntry{key, nullptr, 0, 0});
    return enqueue(st);
  }
  if (hit->state < 0) return enqueue(st);
  if (hit->state == 0) {                                        // second sighting: capture and instantiate
    cudaStream_t& cs = g_ns_capture_stream[key.dev];
    if (!cs && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { cs = nullptr; hit->state = -1; cudaGetLastError(); return enqueue(st); }
    const unsigned long long before = launches();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { hit->state = -1; cudaGetLastError(); return enqueue(st); }
    const int rc = enqueue(cs);
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    const int captured = (int)(launches() - before);
    count_launch(-captured);                                    // nothing has run yet
    cudaGraphExec_t exec = nullptr;
    if (rc != OTK_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      hit->state = -1;
      return enqueue(st);
    }
    cudaGraphDestroy(graph);
    hit->exec = exec;
    hit->launches = captured;
    hit->state = 1;
  }
  OTK_CUDA(cudaGraphLaunch(hit->exec, st));
  count_launch(hit->launches);
  return OTK_OK;
}

// Coupled Newton-Schulz on (A + ridge I)/c:  T = (3I - Z Y)/2, Y <- Y T, Z <- T Z.
// On return w.Y[*cur] ~ sqrt(A/c), w.Z[*cur] ~ (A/c)^-1/2 (unsymmetrised), c in w.c; *verdict says whether the
// residual reached the working-precision floor within the budget.
template <typename W>
static int ns_solve(const void* a, int dt, int64_t L, int64_t d, double ridge, int iters, NsWork<W>& w, int* cur_out,
                    int* verdict, int* used, cudaStream_t st) {
  const int64_t dd = d * d;
  const bool f32 = sizeof(W) == 4;
  bool planes = false;
  if constexpr (sizeof(W) == 4) planes = ns_planes_eligible(d) && L <= 65535;
  OTK_TRY(frob_norm(a, dt, L, d, ridge, f32 ? 0.0 : NS_F64_REL_RIDGE, w.c, w.ridge_l, st));
  if constexpr (sizeof(W) == 4) {
    if (planes) ns_init_planes_kernel<<<ew_grid(L * dd), 256, 0, st>>>(a, dt, L, d, w.ridge_l, w.c, w.Yh[0], w.Yl[0], w.Zh[0], w.Zl[0]);
  }
  if (!planes) ns_init_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(a, dt, L, d, w.ridge_l, w.c, w.Y[0], w.Z[0]);
  OTK_LAUNCH_CHECK();
  OTK_CUDA(cudaMemsetAsync(w.resid, 0, (size_t)NS_MAX_ITERS * L * 8, st));
  const bool adaptive = iters <= 0;
  // accelerated first iterations (adaptive runs only: a caller that fixes the iteration count gets the classical step)
  const int accel = adaptive ? (f32 ? NS_ACCEL_STEPS : g_ns_accel_steps) : 0;
  const int budget = f32 ? NS_F32_MAX_ITERS : NS_F64_MAX_ITERS;
  const int max_iters = adaptive ? budget : (iters < NS_MAX_ITERS ? iters : NS_MAX_ITERS);
  // quadratic convergence: once ||I - ZY||_F^2 < tol_near one more update lands on the floor
  const double tol_done = f32 ? 1e-9 : 1e-22, tol_near = f32 ? 9e-4 : 1e-8;
  // The host enqueues iterations without waiting: the stopping rule runs on the device (ns_ctrl_kernel) and the kernels
  // of iterations past ctrl[0] retire immediately.  One readback after the first NS_FIRST_BATCH iterations, then one per
  // NS_NEXT_BATCH, instead of one host synchronisation per iteration.
  constexpr int NS_FIRST_BATCH = 12, NS_NEXT_BATCH = 4;
  ns_ctrl_init_kernel<<<1, 32, 0, st>>>(w.ctrl, max_iters, adaptive ? NS_SLOW : NS_CONVERGED);
  OTK_LAUNCH_CHECK();
  int h_ctrl[4] = {max_iters, adaptive ? NS_SLOW : NS_CONVERGED, 0, 0};
  int enq = 0;
  while (enq < max_iters) {
    int n = adaptive ? (enq == 0 ? (planes ? NS_FIRST_BATCH : 6) : (planes ? NS_NEXT_BATCH : 1)) : max_iters;
    if (enq + n > max_iters) n = max_iters - enq;
    bool done_planes = false;
    if constexpr (sizeof(W) == 4) {
      if (planes) {
        auto enqueue = [&](cudaStream_t s) -> int {
          for (int k = enq; k < enq + n; ++k) {
            const int cur = k & 1;
            if (k < accel) {
              // accelerated step: M = Z Y (residual of the raw product as usual), then T = c M^2 + b M + a I
              GemmArgs<float> zy = plane_args(w.Zh[cur], w.Zl[cur], w.Yh[cur], w.Yl[cur], w.Mh, w.Ml, d, 1.f, 0.f, w.resid + (size_t)k * L);
              OTK_TRY(plane_gemm2(zy, nullptr, L, w.ctrl, k, s));
              GemmArgs<float> mm = plane_args(w.Mh, w.Ml, w.Mh, w.Ml, w.Th, w.Tl, d, (float)NS_ACC_C, (float)NS_ACC_A, nullptr);
              mm.add = w.Mh; mm.add_lo = w.Ml; mm.add_scale = (float)NS_ACC_B;
              OTK_TRY(plane_gemm2(mm, nullptr, L, w.ctrl, k, s));
            } else {
              GemmArgs<float> zy = plane_args(w.Zh[cur], w.Zl[cur], w.Yh[cur], w.Yl[cur], w.Th, w.Tl, d, -0.5f, 1.5f, w.resid + (size_t)k * L);
              OTK_TRY(plane_gemm2(zy, nullptr, L, w.ctrl, k, s));
            }
            // the stopping rule on the residuals of iteration k rides in the paired launch (one warp of its first CTA):
            // it only has to act before the launches of iteration k + 1
            const NsCtrlEval ev{w.resid + (size_t)k * L, L, k, max_iters, tol_done, tol_near, w.ctrl};
            GemmArgs<float> yt = plane_args(w.Yh[cur], w.Yl[cur], w.Th, w.Tl, w.Yh[cur ^ 1], w.Yl[cur ^ 1], d, 1.f, 0.f, nullptr);
            GemmArgs<float> tz = plane_args(w.Th, w.Tl, w.Zh[cur], w.Zl[cur], w.Zh[cur ^ 1], w.Zl[cur ^ 1], d, 1.f, 0.f, nullptr);
            OTK_TRY(plane_gemm2(yt, &tz, L, w.ctrl, k, s, adaptive ? &ev : nullptr));
          }
          return OTK_OK;
        };
        int dev = 0;
        cudaGetDevice(&dev);
        OTK_TRY(ns_batch(NsGraphKey{w.ctrl, w.Yh[0], L, d, enq, n, max_iters, (adaptive ? 1 : 0) + 2 * accel, dev}, st, enqueue));
        done_planes = true;
      }
    }
    for (int k = enq; k < enq + n && !done_planes; ++k) {
      const int cur = k & 1;
      {
        // generic engines (FFMA / DFMA): the launches are unconditional, so they are only enqueued up to the next readback
        if (k < accel) {
          GemmArgs<W> g = nn_args_t<W>(w.Z[cur], w.Y[cur], w.M, d, dd, W(1));
          g.resid = w.resid + (size_t)k * L;
          OTK_TRY(gemm_any(g, L, st));
          GemmArgs<W> mm = nn_args_t<W>(w.M, w.M, w.T, d, dd, W(NS_ACC_C));
          mm.diag_add = W(NS_ACC_A);
          mm.add = w.M; mm.add_scale = W(NS_ACC_B);
          OTK_TRY(gemm_any(mm, L, st));
        } else {
          GemmArgs<W> g = nn_args_t<W>(w.Z[cur], w.Y[cur], w.T, d, dd, W(-0.5));
          g.diag_add = W(1.5);
          g.resid = w.resid + (size_t)k * L;
          OTK_TRY(gemm_any(g, L, st));
        }
        OTK_TRY(gemm_any(nn_args_t<W>(w.Y[cur], w.T, w.Y[cur ^ 1], d, dd, W(1)), L, st));
        OTK_TRY(gemm_any(nn_args_t<W>(w.T, w.Z[cur], w.Z[cur ^ 1], d, dd, W(1)), L, st));
        if (adaptive) {
          ns_ctrl_kernel<<<1, 256, 0, st>>>(w.resid + (size_t)k * L, L, k, max_iters, tol_done, tol_near, w.ctrl);
          OTK_LAUNCH_CHECK();
        }
      }
    }
    enq += n;
    if (!adaptive) break;
    OTK_CUDA(cudaMemcpyAsync(h_ctrl, w.ctrl, sizeof(h_ctrl), cudaMemcpyDeviceToHost, st));
    OTK_CUDA(cudaStreamSynchronize(st));
    if (h_ctrl[0] <= enq) break;   // converged (or diverged) within what has been enqueued
  }
  // conditional (tcgen05) launches stop at ctrl[0]; the unconditional engines executed everything that was enqueued
  const int executed = planes ? (h_ctrl[0] < enq ? h_ctrl[0] : enq) : enq;
  int cur = executed & 1;
  const int stop_at = h_ctrl[0];
  *verdict = h_ctrl[1];
  if constexpr (sizeof(W) == 4) {
    if (planes) {
      join_planes_kernel<<<ew_grid(L * dd), 256, 0, st>>>(w.Yh[cur], w.Yl[cur], L * dd, w.Y[cur]);
      join_planes_kernel<<<ew_grid(L * dd), 256, 0, st>>>(w.Zh[cur], w.Zl[cur], L * dd, w.Z[cur]);
      count_launch(1);
      OTK_LAUNCH_CHECK();
    }
  }
  *cur_out = cur;
  *used = stop_at < max_iters ? stop_at : max_iters;
  return OTK_OK;
}

template <typename W>
static int sqrtm_impl(const void* a, int64_t L, int64_t dim, int dtype, double ridge, int iters, void* root, void* iroot,
                      void* workspace, size_t workspace_bytes, int* verdict, int* used, cudaStream_t st) {
  Arena ar(workspace, workspace_bytes);
  NsWork<W> w; w.carve(ar, L, dim);
  int cur = 0;
  OTK_TRY(ns_solve<W>(a, dtype, L, dim, ridge, iters, w, &cur, verdict, used, st));
  const W* none = nullptr;
  if (root) {
    sym_scale_kernel<W><<<ew_grid(L * dim * dim), 256, 0, st>>>(w.Y[cur], none, L, dim, w.c, 0.5, 1.0, 0, 0, 0.0, root, dtype);
    OTK_LAUNCH_CHECK();
  }
  if (iroot) {
    sym_scale_kernel<W><<<ew_grid(L * dim * dim), 256, 0, st>>>(w.Z[cur], none, L, dim, w.c, -0.5, 1.0, 0, 0, 0.0, iroot, dtype);
    OTK_LAUNCH_CHECK();
  }
  return OTK_OK;
}

// Shared by w2_gaussian and transport_operator.  Given covariances P (rooted, + ridge) and Q:
//   S  = P^1/2  (first-order ridge correction: sqrt(P) = sqrt(P + eI) - e/2 (P + eI)^-1/2), Zp = (P + eI)^-1/2,
//   leaves sqrt(S Q S)/sqrt(c) in w.Y[*cur2] with c in w.c.
template <typename W>
static int rooted_mix(const void* P, const void* Q, int dt, int64_t L, int64_t d, double ridge, int iters, NsWork<W>& w,
                      W* S, W* Zp, W* Q32, W* G, W* mix, int* cur2, int* verdict, int* used_first, cudaStream_t st,
                      int* used_mix = nullptr) {
  const int64_t dd = d * d;
  const W* none = nullptr;
  int cur = 0, v1 = 0, v2 = 0, used2 = 0;
  OTK_TRY(ns_solve<W>(P, dt, L, d, ridge, iters, w, &cur, &v1, used_first, st));
  const int wdt = sizeof(W) == 8 ? OTK_F64 : OTK_F32;
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(w.Y[cur], w.Z[cur], L, d, w.c, 0.5, 1.0, -0.5, -0.5 * ridge, 0.0, S, wdt);
  OTK_LAUNCH_CHECK();
  if (Zp) {
    sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(w.Z[cur], none, L, d, w.c, -0.5, 1.0, 0, 0, 0.0, Zp, wdt);
    OTK_LAUNCH_CHECK();
  }
  cast_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(Q, dt, L * dd, Q32);
  OTK_LAUNCH_CHECK();
  OTK_TRY(gemm_any(with_scratch(nn_args_t<W>(S, Q32, G, d, dd, W(1)), w.scratch), L, st));     // S Q
  OTK_TRY(gemm_any(with_scratch(nn_args_t<W>(G, S, Q32, d, dd, W(1)), w.scratch), L, st));     // (S Q) S
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(Q32, none, L, d, nullptr, 0, 1.0, 0, 0, 0.0, mix, wdt);
  OTK_LAUNCH_CHECK();
  OTK_TRY(ns_solve<W>(mix, wdt, L, d, 0.0, iters, w, cur2, &v2, &used2, st));
  if (used_mix) *used_mix = used2;
  *verdict = (v1 == NS_CONVERGED && v2 == NS_CONVERGED) ? NS_CONVERGED : NS_SLOW;
  return OTK_OK;
}

template <typename W>
static int w2_impl(const void* mean_s, const void* mean_t, const void* cov_s, const void* cov_t, int64_t L, int64_t dim,
                   int dtype, int iters, double* w2, void* workspace, size_t workspace_bytes, int* verdict,
                   cudaStream_t st) {
  int used_first = 0, used_mix = 0;
  Arena ar(workspace, workspace_bytes);
  NsWork<W> w; w.carve(ar, L, dim);
  const size_t n = (size_t)L * dim * dim;
  W *S = ar.take<W>(n), *Q32 = ar.take<W>(n), *G = ar.take<W>(n), *mix = ar.take<W>(n);
  int cur2 = 0;
  // the reference roots the TARGET covariance for the distance (w2_utils.py:70-71)
  OTK_TRY(rooted_mix<W>(cov_t, cov_s, dtype, L, dim, 0.0, iters, w, S, nullptr, Q32, G, mix, &cur2, verdict, &used_first, st,
                        &used_mix));
  // accuracy gate of the fp32 engine: both roots must have converged within the well-conditioned budget
  if (sizeof(W) == 4 && (used_first > NS_F32_ROOT_ITERS || used_mix > NS_F32_ROOT_ITERS)) *verdict = NS_SLOW;
  w2_trace_kernel<W><<<(unsigned)L, 256, 0, st>>>(mean_s, mean_t, cov_s, cov_t, dtype, dim, w.Y[cur2], w.c, w2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

template <typename W>
static int operator_impl(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype, double pg_star, int iters,
                         void* T, const void* mean_s, const void* mean_t, double* w2, void* workspace,
                         size_t workspace_bytes, int* verdict, int* used_first, cudaStream_t st) {
  Arena ar(workspace, workspace_bytes);
  NsWork<W> w; w.carve(ar, L, dim);
  const size_t n = (size_t)L * dim * dim;
  const int64_t d = dim, dd = dim * dim;
  const W* none = nullptr;
  const int wdt = sizeof(W) == 8 ? OTK_F64 : OTK_F32;
  W *S = ar.take<W>(n), *Zp = ar.take<W>(n), *Q32 = ar.take<W>(n), *G = ar.take<W>(n), *mix = ar.take<W>(n);
  int cur2 = 0;
  // the map roots the SOURCE covariance (w2_utils.py:766-767); the 1e-8 ridge of the inverse root is kept
  OTK_TRY(rooted_mix<W>(cov_s, cov_t, dtype, L, dim, 1e-8, iters, w, S, Zp, Q32, G, mix, &cur2, verdict, used_first, st));
  if (w2) {
    w2_trace_kernel<W><<<(unsigned)L, 256, 0, st>>>(mean_s, mean_t, cov_s, cov_t, dtype, dim, w.Y[cur2], w.c, w2);
    OTK_LAUNCH_CHECK();
  }
  // R = sqrt(mix) = sqrt(c) * sym(Y);  T = (1-p) Zp R Zp + p I
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(w.Y[cur2], none, L, d, w.c, 0.5, 1.0, 0, 0, 0.0, mix, wdt);
  OTK_LAUNCH_CHECK();
  OTK_TRY(gemm_any(with_scratch(nn_args_t<W>(Zp, mix, G, d, dd, W(1)), w.scratch), L, st));
  OTK_TRY(gemm_any(with_scratch(nn_args_t<W>(G, Zp, Q32, d, dd, W(1)), w.scratch), L, st));
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(Q32, none, L, d, nullptr, 0, 1.0 - pg_star, 0, 0, pg_star, T, dtype);
  OTK_LAUNCH_CHECK();
  if constexpr (sizeof(W) == 4) {
    // accuracy gate of the fp32 engine: the Monge map must satisfy T Cs T = Ct.  Z R Z cancels by a factor
    // cond(Cs), so an ill-conditioned source covariance shows up here and sends the computation to fp64.
    if (*verdict == NS_CONVERGED && L <= 256) {
      float* T0 = S;  // S, Zp, mix are free now
      float *Cs32 = Zp, *Ct32 = mix;
      sym_scale_kernel<float><<<ew_grid(L * dd), 256, 0, st>>>(Q32, none, L, d, nullptr, 0, 1.0, 0, 0, 0.0, T0, OTK_F32);
      cast_kernel<float><<<ew_grid(L * dd), 256, 0, st>>>(cov_s, dtype, L * dd, Cs32);
      cast_kernel<float><<<ew_grid(L * dd), 256, 0, st>>>(cov_t, dtype, L * dd, Ct32);
      count_launch(2);
      OTK_LAUNCH_CHECK();
      OTK_TRY(gemm_any(with_scratch(nn_args_t<float>(T0, Cs32, G, d, dd, 1.f), w.scratch), L, st));
      OTK_TRY(gemm_any(with_scratch(nn_args_t<float>(G, T0, Q32, d, dd, 1.f), w.scratch), L, st));
      OTK_CUDA(cudaMemsetAsync(w.resid, 0, (size_t)2 * L * 8, st));
      rel_residual_kernel<<<dim3(32, (unsigned)L), 256, 0, st>>>(Q32, Ct32, d, w.resid);
      OTK_LAUNCH_CHECK();
      double host_rel[512];
      OTK_CUDA(cudaMemcpyAsync(host_rel, w.resid, (size_t)2 * L * 8, cudaMemcpyDeviceToHost, st));
      OTK_CUDA(cudaStreamSynchronize(st));
      for (int64_t l = 0; l < L; ++l) host_rel[l] = sqrt(host_rel[2 * l] / fmax(host_rel[2 * l + 1], 1e-300));
      for (int64_t l = 0; l < L; ++l)
        if (!(host_rel[l] < riccati_tol(d))) *verdict = NS_SLOW;
    }
  }
  return OTK_OK;
}

// =====================================================================================================================
// Fast path of the deterministic operator (the call GaussianTransport.compute() makes, transport/gaussian_transport.py:64-78).
//
// The general path above is a chain of ~110 launches with three host read-backs (iteration counts of the two solves, the
// Riccati residual): at d <= 512 it is bound by the host's launch rate, not by the GPU.  This path runs the same
// arithmetic OPTIMISTICALLY - a fixed budget of device-gated Newton-Schulz iterations per solve, every decision that needs
// an iteration count (which ping-pong buffer holds the result) taken on the device - as ~70 launches with no read-back in
// between, so the whole call is captured once into a CUDA graph (keyed on the operand / workspace pointers) and replayed;
// operands flow between the products as TF32 hi/lo planes (no split / join / cast passes).  One status block is read back
// at the end: if a solve did not converge within the budget, needed more iterations than the fp32 accuracy gates allow,
// or the map fails the Riccati check, the general path (with its fp64 escalation) redoes the call.
// =====================================================================================================================
constexpr int NS_FAST_ITERS = 12;

struct FastStatus { int ctrl1[4]; int ctrl2[4]; };   // followed by acc[2 L] doubles (Riccati numerators / denominators)

__device__ __forceinline__ double ns_scale_dev(const double* csq, const double* cmax, int64_t l) {
  const double fro = sqrt(csq[l]), rs = cmax[l];
  const double c0 = (rs > 0 && rs < fro) ? rs : fro;
  return c0 > 0 ? c0 : 1.0;
}
__device__ __forceinline__ float plane_sym(const float* __restrict__ h, const float* __restrict__ l, int64_t e, int64_t et) {
  return 0.5f * ((h[e] + l[e]) + (h[et] + l[et]));
}
__device__ __forceinline__ void plane_store(float* __restrict__ h, float* __restrict__ l, int64_t e, float v) {
  float a, b;
  ptx::split_tf32(v, a, b);
  h[e] = a; l[e] = b;
}

// row norms of (MODE 0) A + ridge I read from `a`, or (MODE 1) sym(Rh + Rl); also resets the loop control of the solve
template <int MODE>
__global__ void fast_norm_kernel(const void* a, int dt, const float* __restrict__ Rh, const float* __restrict__ Rl, int64_t dim,
                                 double ridge, double* csq, double* cmax, int* ctrl, int max_iters) {
  const int64_t l = blockIdx.y;
  const int lane = threadIdx.x % 32;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { ctrl[0] = max_iters; ctrl[1] = NS_SLOW; ctrl[2] = 0; ctrl[3] = 0; }
  double sq_tot = 0, mx = 0;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; row < dim; row += (int64_t)gridDim.x * (blockDim.x / 32)) {
    double sq = 0, ab = 0;
    for (int64_t j = lane; j < dim; j += 32) {
      const int64_t e = l * dim * dim + row * dim + j;
      double v;
      if (MODE == 0) { v = load_real(a, e, dt); if (j == row) v += ridge; }
      else v = (double)plane_sym(Rh, Rl, e, l * dim * dim + j * dim + row);
      sq += v * v;
      ab += fabs(v);
    }
    sq = warp_sum(sq); ab = warp_sum(ab);
    sq_tot += sq;
    mx = ab > mx ? ab : mx;
  }
  if (lane == 0) {
    atomicAdd(&csq[l], sq_tot);
    atomicMax(reinterpret_cast<unsigned long long*>(&cmax[l]), (unsigned long long)__double_as_longlong(mx));
  }
}

// Y = value / c as planes, Z = I
template <int MODE>
__global__ void fast_init_kernel(const void* a, int dt, const float* __restrict__ Rh, const float* __restrict__ Rl, int64_t L,
                                 int64_t dim, double ridge, const double* csq, const double* cmax, float* Yh, float* Yl,
                                 float* Zh, float* Zl) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    double v;
    if (MODE == 0) v = load_real(a, e, dt) + (i == j ? ridge : 0.0);
    else v = (double)plane_sym(Rh, Rl, e, l * dim * dim + j * dim + i);
    plane_store(Yh, Yl, e, (float)(v / ns_scale_dev(csq, cmax, l)));
    Zh[e] = i == j ? 1.f : 0.f; Zl[e] = 0.f;
  }
}

struct PlanePair { const float *h[2], *l[2]; };   // ping-pong planes of an iterate; the live one is chosen on the device

// after solve 1 (on P + ridge I): S = P^1/2 (first-order ridge correction), Zp = (P + ridge I)^-1/2, Q = cast(other cov)
__global__ void fast_mid_kernel(PlanePair Y, PlanePair Z, const int* ctrl, int budget, const double* csq, const double* cmax,
                                double ridge, const void* q, int dt, int64_t L, int64_t dim, float* Sh, float* Sl, float* Zph,
                                float* Zpl, float* Qh, float* Ql) {
  const int done = ctrl[0] < budget ? ctrl[0] : budget, cur = done & 1;
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), et = l * dim * dim + (r % dim) * dim + r / dim;
    const double c = ns_scale_dev(csq, cmax, l), rc = sqrt(c);
    const double y = plane_sym(Y.h[cur], Y.l[cur], e, et), z = plane_sym(Z.h[cur], Z.l[cur], e, et);
    plane_store(Sh, Sl, e, (float)(y * rc - 0.5 * ridge * z / rc));
    plane_store(Zph, Zpl, e, (float)(z / rc));
    plane_store(Qh, Ql, e, (float)load_real(q, e, dt));
  }
}

// after solve 2: R = mix^1/2 = sqrt(c2) sym(Y2) as planes; W2^2 accumulated from the diagonal (w2 zeroed beforehand)
__global__ void fast_tail_kernel(PlanePair Y, const int* ctrl, int budget, const double* csq, const double* cmax, int64_t L,
                                 int64_t dim, float* Rh, float* Rl, const void* ms, const void* mt, const void* cs,
                                 const void* ct, int dt, double* w2) {
  const int done = ctrl[0] < budget ? ctrl[0] : budget, cur = done & 1;
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim, et = l * dim * dim + j * dim + i;
    const double rc = sqrt(ns_scale_dev(csq, cmax, l));
    const double y = plane_sym(Y.h[cur], Y.l[cur], e, et);
    plane_store(Rh, Rl, e, (float)(y * rc));
    if (w2 && i == j) {
      const double dm = load_real(ms, l * dim + i, dt) - load_real(mt, l * dim + i, dt);
      atomicAdd(&w2[l], dm * dm + load_real(cs, e, dt) + load_real(ct, e, dt) - 2.0 * rc * y);
    }
  }
}

// T = (1-p) sym(Traw) + p I -> caller's dtype; T0 = sym(Traw) and Cs as planes for the Riccati check
__global__ void fast_final_kernel(const float* __restrict__ Th, const float* __restrict__ Tl, int64_t L, int64_t dim, double pg,
                                  void* T, int out_dt, float* T0h, float* T0l, const void* cs, int dt, float* Csh, float* Csl) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    const float t0 = plane_sym(Th, Tl, e, l * dim * dim + j * dim + i);
    store_real(T, e, out_dt, (1.0 - pg) * (double)t0 + (i == j ? pg : 0.0));
    plane_store(T0h, T0l, e, t0);
    plane_store(Csh, Csl, e, (float)load_real(cs, e, dt));
  }
}

// acc[2l] += || (Vh + Vl)_l - target_l ||_F^2 , acc[2l+1] += || target_l ||_F^2
__global__ void fast_residual_kernel(const float* __restrict__ Vh, const float* __restrict__ Vl, const void* target, int dt,
                                     int64_t dim, double* acc) {
  __shared__ double red[32];
  const int64_t l = blockIdx.y;
  double num = 0, den = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < dim * dim; e += (int64_t)gridDim.x * blockDim.x) {
    const double t = load_real(target, l * dim * dim + e, dt), df = (double)(Vh[l * dim * dim + e] + Vl[l * dim * dim + e]) - t;
    num += df * df;
    den += t * t;
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) { atomicAdd(&acc[2 * l], num); atomicAdd(&acc[2 * l + 1], den); }
}

// NS_FAST_ITERS device-gated iterations on the planes of `w` (the enqueue body of ns_solve, without its read-backs)
static int fast_iterations(NsWork<float>& w, int64_t L, int64_t d, int* ctrl, double* resid, cudaStream_t s) {
  const double tol_done = 1e-9, tol_near = 9e-4;
  for (int k = 0; k < NS_FAST_ITERS; ++k) {
    const int cur = k & 1;
    if (k < NS_ACCEL_STEPS) {
      GemmArgs<float> zy = plane_args(w.Zh[cur], w.Zl[cur], w.Yh[cur], w.Yl[cur], w.Mh, w.Ml, d, 1.f, 0.f, resid + (size_t)k * L);
      OTK_TRY(plane_gemm2(zy, nullptr, L, ctrl, k, s));
      GemmArgs<float> mm = plane_args(w.Mh, w.Ml, w.Mh, w.Ml, w.Th, w.Tl, d, (float)NS_ACC_C, (float)NS_ACC_A, nullptr);
      mm.add = w.Mh; mm.add_lo = w.Ml; mm.add_scale = (float)NS_ACC_B;
      OTK_TRY(plane_gemm2(mm, nullptr, L, ctrl, k, s));
    } else {
      GemmArgs<float> zy = plane_args(w.Zh[cur], w.Zl[cur], w.Yh[cur], w.Yl[cur], w.Th, w.Tl, d, -0.5f, 1.5f, resid + (size_t)k * L);
      OTK_TRY(plane_gemm2(zy, nullptr, L, ctrl, k, s));
    }
    const NsCtrlEval ev{resid + (size_t)k * L, L, k, NS_F32_MAX_ITERS, tol_done, tol_near, ctrl};
    GemmArgs<float> yt = plane_args(w.Yh[cur], w.Yl[cur], w.Th, w.Tl, w.Yh[cur ^ 1], w.Yl[cur ^ 1], d, 1.f, 0.f, nullptr);
    GemmArgs<float> tz = plane_args(w.Th, w.Tl, w.Zh[cur], w.Zl[cur], w.Zh[cur ^ 1], w.Zl[cur ^ 1], d, 1.f, 0.f, nullptr);
    OTK_TRY(plane_gemm2(yt, &tz, L, ctrl, k, s, &ev));
  }
  return OTK_OK;
}

struct FastOpArgs {
  const void *cov_s, *cov_t, *mean_s, *mean_t;
  void* T;
  double* w2;
  void* workspace;
  size_t workspace_bytes;
  int64_t L, d;
  int dtype;
  double pg_star;
};
struct FastOpLayout { NsWork<float> w; float* pl[8]; double *c1, *c2, *resid2; FastStatus* status; double* acc; size_t zero_lo, zero_hi; };

static bool fast_layout(const FastOpArgs& a, FastOpLayout* o) {
  Arena ar(a.workspace, a.workspace_bytes);
  o->w.carve(ar, a.L, a.d);
  const size_t n = (size_t)a.L * a.d * a.d;
  for (int i = 0; i < 8; ++i) o->pl[i] = ar.take<float>(n);
  // one zero-filled region: [c1 sq | c1 max | c2 sq | c2 max | resid1 | resid2 | status | acc]
  ar.off = align_up(ar.off, 256);
  o->zero_lo = ar.off;
  o->c1 = ar.take<double>(2 * (size_t)a.L);
  o->c2 = ar.take<double>(2 * (size_t)a.L);
  o->resid2 = ar.take<double>((size_t)NS_MAX_ITERS * a.L);
  o->status = ar.take<FastStatus>(1);
  o->acc = ar.take<double>(2 * (size_t)a.L);
  o->zero_hi = ar.off;
  return ar.ok();
}

static int fast_enqueue(const FastOpArgs& a, FastOpLayout& o, cudaStream_t st) {
  const int64_t L = a.L, d = a.d, dd = d * d;
  NsWork<float>& w = o.w;
  float *Sh = o.pl[0], *Sl = o.pl[1], *Zph = o.pl[2], *Zpl = o.pl[3], *Qh = o.pl[4], *Ql = o.pl[5], *Gh = o.pl[6], *Gl = o.pl[7];
  const double ridge = 1e-8;    // the reference's ridge of the inverse root (w2_utils.py:766)
  OTK_CUDA(cudaMemsetAsync(static_cast<char*>(a.workspace) + o.zero_lo, 0, o.zero_hi - o.zero_lo, st));
  OTK_CUDA(cudaMemsetAsync(w.resid, 0, (size_t)NS_MAX_ITERS * L * 8, st));
  if (a.w2) OTK_CUDA(cudaMemsetAsync(a.w2, 0, (size_t)L * 8, st));
  int64_t nb = ceil_div(d, 8);
  if (nb > 128) nb = 128;
  const dim3 ngrid((unsigned)nb, (unsigned)L);
  const PlanePair Y{{w.Yh[0], w.Yh[1]}, {w.Yl[0], w.Yl[1]}}, Z{{w.Zh[0], w.Zh[1]}, {w.Zl[0], w.Zl[1]}};
  auto mul = [&](const float* Ah, const float* Al, const float* Bh, const float* Bl, float* Ch, float* Cl) {
    return plane_gemm2(plane_args(Ah, Al, Bh, Bl, Ch, Cl, d, 1.f, 0.f, nullptr), nullptr, L, nullptr, 0, st);
  };
  // ---- solve 1: roots of the SOURCE covariance (+ ridge)
  fast_norm_kernel<0><<<ngrid, 256, 0, st>>>(a.cov_s, a.dtype, nullptr, nullptr, d, ridge, o.c1, o.c1 + L, o.status->ctrl1, NS_FAST_ITERS);
  fast_init_kernel<0><<<ew_grid(L * dd), 256, 0, st>>>(a.cov_s, a.dtype, nullptr, nullptr, L, d, ridge, o.c1, o.c1 + L, w.Yh[0], w.Yl[0],
                                                       w.Zh[0], w.Zl[0]);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  OTK_TRY(fast_iterations(w, L, d, o.status->ctrl1, w.resid, st));
  fast_mid_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Y, Z, o.status->ctrl1, NS_FAST_ITERS, o.c1, o.c1 + L, ridge, a.cov_t, a.dtype, L, d,
                                                  Sh, Sl, Zph, Zpl, Qh, Ql);
  OTK_LAUNCH_CHECK();
  OTK_TRY(mul(Sh, Sl, Qh, Ql, Gh, Gl));            // S Ct
  OTK_TRY(mul(Gh, Gl, Sh, Sl, Qh, Ql));            // (S Ct) S  -> Q planes
  // ---- solve 2: root of sym(S Ct S)
  fast_norm_kernel<1><<<ngrid, 256, 0, st>>>(nullptr, 0, Qh, Ql, d, 0.0, o.c2, o.c2 + L, o.status->ctrl2, NS_FAST_ITERS);
  fast_init_kernel<1><<<ew_grid(L * dd), 256, 0, st>>>(nullptr, 0, Qh, Ql, L, d, 0.0, o.c2, o.c2 + L, w.Yh[0], w.Yl[0], w.Zh[0], w.Zl[0]);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  OTK_TRY(fast_iterations(w, L, d, o.status->ctrl2, o.resid2, st));
  fast_tail_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Y, o.status->ctrl2, NS_FAST_ITERS, o.c2, o.c2 + L, L, d, Gh, Gl, a.mean_s, a.mean_t,
                                                   a.cov_s, a.cov_t, a.dtype, a.w2);
  OTK_LAUNCH_CHECK();
  OTK_TRY(mul(Zph, Zpl, Gh, Gl, Qh, Ql));          // Zp R
  OTK_TRY(mul(Qh, Ql, Zph, Zpl, Sh, Sl));          // (Zp R) Zp = Traw  -> S planes
  fast_final_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Sh, Sl, L, d, a.pg_star, a.T, a.dtype, Gh, Gl, a.cov_s, a.dtype, Qh, Ql);
  OTK_LAUNCH_CHECK();
  // ---- Riccati check of the pg_star = 0 map: T0 Cs T0 = Ct
  OTK_TRY(mul(Gh, Gl, Qh, Ql, Sh, Sl));            // T0 Cs
  OTK_TRY(mul(Sh, Sl, Gh, Gl, Zph, Zpl));          // (T0 Cs) T0
  fast_residual_kernel<<<dim3(32, (unsigned)L), 256, 0, st>>>(Zph, Zpl, a.cov_t, a.dtype, d, o.acc);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// graph cache of the fast path (same life cycle as the Newton-Schulz batch cache: plain launches on the first sighting of
// a key, capture on the second, replay afterwards)
struct FastGraphKey {
  const void* p[7];
  int64_t L, d;
  int dtype, dev;
  double pg;
  bool operator==(const FastGraphKey& o) const {
    for (int i = 0; i < 7; ++i) if (p[i] != o.p[i]) return false;
    return L == o.L && d == o.d && dtype == o.dtype && dev == o.dev && pg == o.pg;
  }
};
struct FastGraphEntry { FastGraphKey key; cudaGraphExec_t exec; int launches; int state; };
static std::vector<FastGraphEntry> g_fast_graphs;
int g_fast_counters[4] = {0, 0, 0, 0};   // tuning aid: accepted, rejected, graph replays, eager runs (otkdbg_fast_counters)
constexpr size_t FAST_GRAPH_CACHE = 16;

// 1 = done (T, w2 written), 0 = not taken / not accepted (the caller runs the general path), < 0 = error
static int operator_fast(const FastOpArgs& a, cudaStream_t st) {
  if (!ns_planes_eligible(a.d) || a.L > 4096) return 0;
  static const bool fast_on = [] { const char* e = getenv("OTK_OPERATOR_FAST"); return !(e && e[0] == '0'); }();   // tuning aid
  if (!fast_on) return 0;
  FastOpLayout lay;
  if (!fast_layout(a, &lay)) return 0;
  auto enqueue = [&](cudaStream_t s) -> int { return fast_enqueue(a, lay, s); };
  int dev = 0;
  cudaGetDevice(&dev);
  static const bool graphs_on = [] { const char* e = getenv("OTK_NS_GRAPHS"); return !(e && e[0] == '0'); }();
  bool launched = false;
  if (graphs_on && dev >= 0 && dev < 64) {
    const FastGraphKey key{{a.cov_s, a.cov_t, a.mean_s, a.mean_t, a.T, a.w2, a.workspace}, a.L, a.d, a.dtype, dev, a.pg_star};
    std::lock_guard<std::mutex> lock(g_ns_graph_mu);
    FastGraphEntry* hit = nullptr;
    for (auto& e : g_fast_graphs) if (e.key == key) { hit = &e; break; }
    if (!hit) {
      if (g_fast_graphs.size() >= FAST_GRAPH_CACHE) {
        if (g_fast_graphs.front().exec) cudaGraphExecDestroy(g_fast_graphs.front().exec);
        g_fast_graphs.erase(g_fast_graphs.begin());
      }
      g_fast_graphs.push_back(FastGraphEntry{key, nullptr, 0, 0});
    } else if (hit->state == 0) {
      cudaStream_t& cs = g_ns_capture_stream[dev];
      if (!cs && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { cs = nullptr; hit->state = -1; cudaGetLastError(); }
      if (cs && cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const unsigned long long before = launches();
        cudaGraph_t graph = nullptr;
        const int rc = enqueue(cs);
        const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        const int captured = (int)(launches() - before);
        count_launch(-captured);
        cudaGraphExec_t exec = nullptr;
        if (rc == OTK_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
          hit->exec = exec; hit->launches = captured; hit->state = 1;
        } else {
          hit->state = -1;
          cudaGetLastError();
        }
        if (graph) cudaGraphDestroy(graph);
      } else if (cs) {
        hit->state = -1;
        cudaGetLastError();
      }
    }
    if (hit && hit->state == 1) {
      OTK_CUDA(cudaGraphLaunch(hit->exec, st));
      count_launch(hit->launches);
      launched = true;
      ++g_fast_counters[2];
    }
  }
  if (!launched) { OTK_TRY(enqueue(st)); ++g_fast_counters[3]; }
  // ---- the one read-back: loop control of both solves + Riccati sums
  // (the arena aligns every array to 256 bytes: the span from the status block to the end of the sums is what is copied)
  const size_t span = reinterpret_cast<char*>(lay.acc + 2 * a.L) - reinterpret_cast<char*>(lay.status);
  std::vector<char> host(span);
  OTK_CUDA(cudaMemcpyAsync(host.data(), lay.status, span, cudaMemcpyDeviceToHost, st));
  OTK_CUDA(cudaStreamSynchronize(st));
  const FastStatus* hs = reinterpret_cast<const FastStatus*>(host.data());
  const double* acc = reinterpret_cast<const double*>(host.data() + (reinterpret_cast<char*>(lay.acc) - reinterpret_cast<char*>(lay.status)));
  const bool conv = hs->ctrl1[1] == NS_CONVERGED && hs->ctrl2[1] == NS_CONVERGED && hs->ctrl1[2] == 0 && hs->ctrl2[2] == 0 &&
                    hs->ctrl1[0] <= NS_FAST_ITERS && hs->ctrl2[0] <= NS_FAST_ITERS && hs->ctrl1[0] <= NS_F32_OPERATOR_ITERS;
  static const bool dbg = [] { const char* e = getenv("OTK_FAST_DEBUG"); return e && e[0] == '1'; }();
  if (dbg) {
    fprintf(stderr, "operator_fast d=%lld L=%lld: solve1 iters %d verdict %d div %d | solve2 iters %d verdict %d div %d | riccati",
            (long long)a.d, (long long)a.L, hs->ctrl1[0], hs->ctrl1[1], hs->ctrl1[2], hs->ctrl2[0], hs->ctrl2[1], hs->ctrl2[2]);
    for (int64_t l = 0; l < a.L && l < 4; ++l) fprintf(stderr, " %.2e", sqrt(acc[2 * l] / fmax(acc[2 * l + 1], 1e-300)));
    fprintf(stderr, "\n");
  }
  if (!conv) { ++g_fast_counters[1]; return 0; }
  for (int64_t l = 0; l < a.L; ++l) {
    const double rel = sqrt(acc[2 * l] / fmax(acc[2 * l + 1], 1e-300));
    if (!(rel < riccati_tol(a.d))) { ++g_fast_counters[1]; return 0; }
  }
  ++g_fast_counters[0];
  return 1;
}

// out = s0 * in + diag_add * I   (no symmetrisation: the stochastic operator is not symmetric), cast to out_dt
template <typename W>
__global__ void scale_add_kernel(const W* in, int64_t L, int64_t dim, double s0, double diag_add, void* out, int out_dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % (dim * dim);
    store_real(out, e, out_dt, s0 * (double)in[e] + ((r / dim == r % dim) ? diag_add : 0.0));
  }
}

// Stochastic operator, eq. 19 of Freirich et al. (reference _compute_transport_full_mat_stochastic, w2_utils.py:774-793):
//   St = Ct^1/2, iSt = (Ct + 1e-8 I)^-1/2, R = (St Cs St)^1/2, P = Cs^+ (= Cs^-1: the source must be PD here - the
//   reference's own pipeline is non-finite for a rank-deficient source, tests/golden/make_golden.py note),
//   T* = iSt R iSt,  T = (1-p) St R iSt P + p I,  Cw = sqrt(1-p) St (I - St T* P T* St) St.
// Three Newton-Schulz solves and fourteen d x d products, all on the device (the reference: 5 eigh + pinv (SVD) + 14 matmuls).
template <typename W>
static int stochastic_impl(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype, double pg_star, int iters,
                           void* T, void* Cw, void* workspace, size_t workspace_bytes, int* verdict, cudaStream_t st) {
  Arena ar(workspace, workspace_bytes);
  NsWork<W> w; w.carve(ar, L, dim);
  const size_t n = (size_t)L * dim * dim;
  const int64_t d = dim, dd = dim * dim;
  const W* none = nullptr;
  const int wdt = sizeof(W) == 8 ? OTK_F64 : OTK_F32;
  W *St = ar.take<W>(n), *iSt = ar.take<W>(n), *Q32 = ar.take<W>(n), *G = ar.take<W>(n), *mix = ar.take<W>(n);
  W *R = ar.take<W>(n), *P = ar.take<W>(n), *Tst = ar.take<W>(n), *H = ar.take<W>(n);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  int cur = 0, cur2 = 0, v0 = 0, v1 = 0, used = 0;
  auto mul = [&](const W* A, const W* B, W* C) { return gemm_any(with_scratch(nn_args_t<W>(A, B, C, d, dd, W(1)), w.scratch), L, st); };
  // P = Cs^-1 from the inverse root of the source
  OTK_TRY(ns_solve<W>(cov_s, dtype, L, d, 0.0, iters, w, &cur, &v0, &used, st));
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(w.Z[cur], none, L, d, w.c, -0.5, 1.0, 0, 0, 0.0, G, wdt);
  OTK_LAUNCH_CHECK();
  OTK_TRY(mul(G, G, P));
  // roots of the TARGET (roles swapped on purpose, reference :786-787), then R = (St Cs St)^1/2
  OTK_TRY(rooted_mix<W>(cov_t, cov_s, dtype, L, d, 1e-8, iters, w, St, iSt, Q32, G, mix, &cur2, &v1, &used, st));
  sym_scale_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(w.Y[cur2], none, L, d, w.c, 0.5, 1.0, 0, 0, 0.0, R, wdt);
  OTK_LAUNCH_CHECK();
  *verdict = (v0 == NS_CONVERGED && v1 == NS_CONVERGED) ? NS_CONVERGED : NS_SLOW;
  // T = (1-p) St R iSt P + p I
  OTK_TRY(mul(St, R, G));
  OTK_TRY(mul(G, iSt, H));
  OTK_TRY(mul(H, P, G));
  scale_add_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(G, L, d, 1.0 - pg_star, pg_star, T, dtype);
  OTK_LAUNCH_CHECK();
  // T* = iSt R iSt ;  Cw = sqrt(1-p) St (I - St T* P T* St) St
  OTK_TRY(mul(iSt, R, G));
  OTK_TRY(mul(G, iSt, Tst));
  OTK_TRY(mul(St, Tst, G));
  OTK_TRY(mul(G, P, H));
  OTK_TRY(mul(H, Tst, G));
  OTK_TRY(mul(G, St, H));
  scale_add_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(H, L, d, -1.0, 1.0, G, wdt);     // I - St T* P T* St
  OTK_LAUNCH_CHECK();
  OTK_TRY(mul(St, G, H));
  OTK_TRY(mul(H, St, G));
  scale_add_kernel<W><<<ew_grid(L * dd), 256, 0, st>>>(G, L, d, sqrt(1.0 - pg_star), 0.0, Cw, dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

}  // namespace otk
using namespace otk;

extern "C" size_t otk_transport_operator_stochastic_workspace_bytes(int64_t L, int64_t dim) {
  return ns_work_bytes(L, dim) + 9 * align_up((size_t)L * dim * dim * 8, 256) + 4096;
}

extern "C" int otk_transport_operator_stochastic(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype,
                                                 double pg_star, int iters, int polish, void* T, void* Cw, void* workspace,
                                                 size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(cov_s && cov_t && T && Cw && L > 0 && dim > 0, "transport_operator_stochastic: bad arguments");
  OTK_REQUIRE(pg_star >= 0.0 && pg_star <= 1.0, "transport_operator_stochastic: pg_star outside [0, 1]");
  if (!workspace || workspace_bytes < otk_transport_operator_stochastic_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int verdict = NS_SLOW;
  // Cw = St (I - ...) St cancels to zero for a PD source: only the fp64 engine resolves it for fp64 data
  if (polish == 0 && dtype == OTK_F64 && iters <= 0) polish = 1;
  if (polish <= 0) {
    OTK_TRY(stochastic_impl<float>(cov_s, cov_t, L, dim, dtype, pg_star, iters, T, Cw, workspace, workspace_bytes, &verdict, st));
    if (verdict == NS_CONVERGED || polish < 0 || iters > 0) return OTK_OK;
  }
  OTK_TRY(stochastic_impl<double>(cov_s, cov_t, L, dim, dtype, pg_star, iters, T, Cw, workspace, workspace_bytes, &verdict, st));
  if (verdict != NS_CONVERGED && iters <= 0) {
    set_last_error_msg("transport_operator_stochastic: Newton-Schulz did not converge (a covariance is not positive definite)");
    return OTK_ERR_NOT_CONVERGED;
  }
  return OTK_OK;
}

// workspaces are sized for the fp64 escalation
extern "C" size_t otk_sqrtm_workspace_bytes(int64_t L, int64_t dim) { return ns_work_bytes(L, dim) + 4096; }
extern "C" size_t otk_w2_gaussian_workspace_bytes(int64_t L, int64_t dim) {
  return ns_work_bytes(L, dim) + 5 * align_up((size_t)L * dim * dim * 8, 256) + 4096;
}
extern "C" size_t otk_transport_operator_workspace_bytes(int64_t L, int64_t dim) {
  return ns_work_bytes(L, dim) + 6 * align_up((size_t)L * dim * dim * 8, 256) + 4096;
}

// `polish`: 0 = precision policy above (fp32 engine, fp64 escalation for fp64 data); 1 = force the fp64 engine;
//           -1 = fp32 engine only.
extern "C" int otk_sqrtm(const void* a, int64_t L, int64_t dim, int dtype, double ridge, int iters, int polish, void* root,
                         void* iroot, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(a && L > 0 && dim > 0 && (root || iroot), "sqrtm: bad arguments");
  if (!workspace || workspace_bytes < otk_sqrtm_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int verdict = NS_SLOW, used = 0;
  if (polish == 0 && dtype == OTK_F64 && dim <= NS_SMALL_DIM && iters <= 0) polish = 1;
  if (polish <= 0) {
    OTK_TRY(sqrtm_impl<float>(a, L, dim, dtype, ridge, iters, root, iroot, workspace, workspace_bytes, &verdict, &used, st));
    // the inverse root loses ~1e-7 * cond: escalate it earlier than the root
    const bool accurate = verdict == NS_CONVERGED && used <= (iroot ? NS_F32_IROOT_ITERS : NS_F32_ROOT_ITERS);
    if (accurate || polish < 0 || iters > 0) return OTK_OK;
  }
  g_ns_accel_steps = polish <= 0 ? NS_ACCEL_STEPS_ESCALATED : NS_ACCEL_STEPS;   // escalated: the input is ill-conditioned
  const int rc64 = sqrtm_impl<double>(a, L, dim, dtype, ridge, iters, root, iroot, workspace, workspace_bytes, &verdict, &used, st);
  g_ns_accel_steps = NS_ACCEL_STEPS;
  OTK_TRY(rc64);
  if (verdict != NS_CONVERGED && iters <= 0) {
    set_last_error_msg("sqrtm: Newton-Schulz did not converge (the matrix is not positive definite)");
    return OTK_ERR_NOT_CONVERGED;
  }
  return OTK_OK;
}

extern "C" int otk_w2_gaussian(const void* mean_s, const void* mean_t, const void* cov_s, const void* cov_t, int64_t L,
                               int64_t dim, int dtype, int iters, int polish, double* w2, void* workspace,
                               size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(mean_s && mean_t && cov_s && cov_t && w2 && L > 0 && dim > 0, "w2_gaussian: bad arguments");
  if (!workspace || workspace_bytes < otk_w2_gaussian_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int verdict = NS_SLOW;
  if (polish == 0 && dtype == OTK_F64 && dim <= NS_SMALL_DIM && iters <= 0) polish = 1;
  if (polish <= 0) {
    OTK_TRY(w2_impl<float>(mean_s, mean_t, cov_s, cov_t, L, dim, dtype, iters, w2, workspace, workspace_bytes, &verdict, st));
    if (verdict == NS_CONVERGED || polish < 0 || iters > 0) return OTK_OK;
  }
  g_ns_accel_steps = polish <= 0 ? NS_ACCEL_STEPS_ESCALATED : NS_ACCEL_STEPS;
  const int rc64 = w2_impl<double>(mean_s, mean_t, cov_s, cov_t, L, dim, dtype, iters, w2, workspace, workspace_bytes, &verdict, st);
  g_ns_accel_steps = NS_ACCEL_STEPS;
  OTK_TRY(rc64);
  if (verdict != NS_CONVERGED && iters <= 0) {
    set_last_error_msg("w2_gaussian: Newton-Schulz did not converge (a covariance is not positive definite)");
    return OTK_ERR_NOT_CONVERGED;
  }
  return OTK_OK;
}

extern "C" int otk_transport_operator(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype,
                                      double pg_star, int iters, int polish, void* T, const void* mean_s,
                                      const void* mean_t, double* w2, void* workspace, size_t workspace_bytes,
                                      otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(cov_s && cov_t && T && L > 0 && dim > 0, "transport_operator: bad arguments");
  OTK_REQUIRE(!w2 || (mean_s && mean_t), "transport_operator: w2 requested without means");
  if (!workspace || workspace_bytes < otk_transport_operator_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int verdict = NS_SLOW, used = 0;
  if (polish == 0 && dtype == OTK_F64 && dim <= NS_SMALL_DIM && iters <= 0) polish = 1;
  if (polish <= 0 && iters <= 0) {
    // optimistic single-graph path; anything it does not accept is redone below by the general path
    const FastOpArgs fa{cov_s, cov_t, mean_s, mean_t, T, w2, workspace, workspace_bytes, L, dim, dtype, pg_star};
    const int fr = operator_fast(fa, st);
    if (fr < 0) return fr;
    if (fr == 1) return OTK_OK;
  }
  if (polish <= 0) {
    OTK_TRY(operator_impl<float>(cov_s, cov_t, L, dim, dtype, pg_star, iters, T, mean_s, mean_t, w2, workspace,
                                 workspace_bytes, &verdict, &used, st));
    // T = Zp R Zp cancels by a factor cond(Cs): fp32 is only kept while the source root converged quickly
    if ((verdict == NS_CONVERGED && used <= NS_F32_OPERATOR_ITERS) || polish < 0 || iters > 0) return OTK_OK;
  }
  g_ns_accel_steps = polish <= 0 ? NS_ACCEL_STEPS_ESCALATED : NS_ACCEL_STEPS;
  const int rc64 = operator_impl<double>(cov_s, cov_t, L, dim, dtype, pg_star, iters, T, mean_s, mean_t, w2, workspace,
                                         workspace_bytes, &verdict, &used, st);
  g_ns_accel_steps = NS_ACCEL_STEPS;
  OTK_TRY(rc64);
  if (verdict != NS_CONVERGED && iters <= 0) {
    set_last_error_msg("transport_operator: Newton-Schulz did not converge (a covariance is not positive definite)");
    return OTK_ERR_NOT_CONVERGED;
  }
  return OTK_OK;
}
