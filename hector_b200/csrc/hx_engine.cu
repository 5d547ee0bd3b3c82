/* hx_engine.cu -- host side of libhector_b200.so: the engine object behind the C ABI
 * (include/hector_b200.h).  Owns the device buffers, flattens scenarios and parameters into
 * the SoA layout of hx_layout.h, precomputes the member-independent gas series on the host
 * (N2O concentration, halocarbon forcings) and drives the kernels of hx_kernels.cu.
 *
 * There is no CPU fallback: without a CUDA device hx_create fails.
 */
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> /* header-only NVTX 3: ranges show up in ncu / Nsight timelines, no-ops otherwise */

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hector_b200.h"
#include "hx_kernels.h"
#include "hx_names.h"

namespace {

/* one NVTX range per C-ABI call that launches work (tracing) */
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

thread_local std::string g_create_error;

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      char buf_[512];                                                                    \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                      \
      return fail(HX_ERR_CUDA, buf_);                                                    \
    }                                                                                    \
  } while (0)

/* ---- small utility kernels ---- */
/* field `f` of a CTA-tiled array (hx_layout.h) for every device member: A[f][m] = v */
__global__ void k_fill_field(double *A, int f, int nfields, double v, int Mpad) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < Mpad) A[HX_TILED(f, m, nfields)] = v;
}
/* A[f][dev_of_api[i]] = src[i]  (src in API member order) */
__global__ void k_scatter_field(double *A, int f, int nfields, const double *src,
                                const int32_t *dev_of_api, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[HX_TILED(f, dev_of_api[i], nfields)] = src[i];
}
__global__ void k_gather_field(double *dst, const double *A, int f, int nfields,
                               const int32_t *dev_of_api, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = A[HX_TILED(f, dev_of_api[i], nfields)];
}
/* out[m_api][k] = src[yidx[k]][col(m_api)]  (tile transpose through shared memory); col = the
 * member itself for blocks kept in API order (the recorded outputs), dev_of_api[m] for blocks in
 * device order (the tracking maps) */
__global__ void k_fetch_transpose(double *out, const double *src, const int32_t *yidx,
                                  const int32_t *dev_of_api, int n_dates, int M, size_t Mpad) {
  __shared__ double tile[32][33];
  const int m0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, m = m0 + threadIdx.x;
    if (k < n_dates && m < M)
      tile[j][threadIdx.x] = src[(size_t)yidx[k] * Mpad + (dev_of_api ? dev_of_api[m] : m)];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int m = m0 + j, k = k0 + threadIdx.x;
    if (k < n_dates && m < M) out[(size_t)m * n_dates + k] = tile[threadIdx.x][j];
  }
}
/* out[m_api][k] = src[k][dev_of_api[m_api]] for 32-bit words (tracking key masks) */
__global__ void k_fetch_masks(uint32_t *out, const uint32_t *src, const int32_t *dev_of_api, int nk,
                              int M, size_t Mpad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * nk) return;
  const int m = i / nk, k = i % nk;
  out[i] = src[(size_t)k * Mpad + dev_of_api[m]];
}
/* live tracking maps of the first HX_NPOOL slots: out[m_api][q] = T[tile][q][lane] */
__global__ void k_gather_track(double *out, uint32_t *out_mask, const double *T, const uint32_t *TK,
                               const int32_t *dev_of_api, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = HX_NPOOL * HX_NSRC;
  if (i >= M * nq) return;
  const int m = i / nq, q = i % nq;
  const int dm = dev_of_api[m];
  out[i] = T[HX_TILED(q, dm, TS_COUNT * HX_NSRC)];
  if (q < HX_NPOOL) out_mask[m * HX_NPOOL + q] = TK[HX_TILED(q, dm, TS_COUNT)];
}
/* copy member `src_m`'s state column into every active member (shared spin-up, E-7) */
__global__ void k_broadcast_state(double *S, int32_t *spinup_steps, int32_t *status,
                                  int32_t *fail_year, int src_m, int n_state, size_t Mpad) {
  size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (m >= Mpad || status[m] < 0 || (int)m == src_m) return;
  for (int i = 0; i < n_state; ++i)
    S[HX_TILED(i, m, SI_COUNT)] = S[HX_TILED(i, src_m, SI_COUNT)];
  spinup_steps[m] = spinup_steps[src_m];
  status[m] = status[src_m];
  fail_year[m] = fail_year[src_m];
}

struct Engine {
  hx_config cfg{};
  HxConst C{};
  HxDev d{};
  std::string err;
  int M = 0, Mpad = 0, nrow = 0, nscen = 0;
  bool prepared = false, params_dirty = false;
  int cur_row = 0;
  /* earliest table row changed since the last set-up (dated series edits; 0 for parameters):
   * the state of every year before it is unaffected, which is what lets hx_reset_date serve R's
   * setvar(core, dates, ...) + reset(core, min(dates) - 1) */
  int dirty_from_row = 0x7fffffff;
  /* something the spin-up or the initial state depends on changed since the post-spin-up
   * snapshot was taken (kSpinupParams, M0, biome inputs, any series): otherwise a reset after a
   * parameter change restores the snapshot and only redoes the DOECLIM set-up (E-7: the spin-up
   * is computed once per distinct spin-up parameter set) */
  bool spinup_dirty = true;
  static bool affects_spinup(int pi) {
    if (pi == PI_M0) return true; /* initial CH4 */
    for (int k : hx::kSpinupParams)
      if (k == pi) return true;
    return false;
  }
  double last_run_ms = 0.0;

  /* host-side inputs */
  std::vector<std::vector<double>> raw;      /* [scen][RAW_COUNT * nrow], series-major */
  std::vector<std::vector<char>> raw_set;    /* [scen][RAW_COUNT] */
  /* user constraints, [scen][CN_COUNT * nrow] series-major, NaN = no entry for that year */
  std::vector<std::vector<double>> cons;
  bool tables_dirty = false;                 /* a series changed after hx_prepare */
  bool tables_constrained = false;           /* some scenario carries a CO2/NBP/CH4/RF_tot/tas constraint */
  bool tables_nbp = false;                   /* ... an NBP constraint */
  std::vector<int32_t> member_scen;          /* API order */
  double pscalar[PI_COUNT];
  std::vector<double> pvec[PI_COUNT];        /* per-member overrides (API order), host copy */
  bool pvec_on_device_only[PI_COUNT];
  double baseyear = 1750, UC_N2O = 4.8, TN2O0 = 132;
  int max_spinup = 2000;
  int tracking_date = 9999, track_every = 1; /* [core] trackingDate; recording stride in years */
  std::vector<int> track_years;              /* recorded years, slot order */
  double halo_tau[HX_NHALO], halo_rho[HX_NHALO], halo_delta[HX_NHALO], halo_H0[HX_NHALO],
      halo_mm[HX_NHALO];
  std::vector<int> out_sel;                  /* OUT_* ids in slot order */

  /* permutation API <-> device */
  std::vector<int32_t> dev_of_api;
  int32_t *d_dev_of_api = nullptr, *d_api_of_dev = nullptr;

  /* device buffers */
  double *d_P = nullptr, *d_S = nullptr, *d_S_snap = nullptr, *d_D = nullptr, *d_ker = nullptr, *d_conv = nullptr,
         *d_sst = nullptr, *d_tland = nullptr, *d_out = nullptr, *d_X = nullptr, *d_scen = nullptr,
         *d_stage = nullptr;
  int32_t *d_block_scen = nullptr, *d_status = nullptr, *d_status_snap = nullptr,
          *d_status_post = nullptr,
          *d_fail_year = nullptr, *d_spinup_steps = nullptr, *d_yidx = nullptr;
  unsigned long long *d_counters = nullptr;
  unsigned *d_sched = nullptr;
  double *d_T = nullptr, *d_TO = nullptr, *d_REC = nullptr;
  unsigned char *d_YCNT = nullptr;
  uint32_t *d_TK = nullptr, *d_TOK = nullptr;
  int32_t *d_trk_fail = nullptr; /* [Mpad] year in which the replay saw a bad mix (0: none) */
  size_t stage_bytes = 0, yidx_cap = 0;
  double *h_pinned = nullptr;
  size_t pinned_bytes = 0;

  cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_seg = nullptr, ev_copy = nullptr;
  cudaStream_t setup_stream = nullptr; /* DOECLIM set-up beside the shared spin-up */
  cudaEvent_t ev_setup_a = nullptr, ev_setup_b = nullptr;
  /* multi-GPU exchange over peer memory: peers' output blocks opened through CUDA IPC */
  std::vector<double *> peer_out;
  std::vector<cudaStream_t> peer_stream; /* one per peer: the pulls spread over the copy engines */
  int peer_self = -1;
  cudaEvent_t ev_user[16] = {};
  /* push exchange (hx_xchg_*): this rank's gather block [peers][n_out][nyears][Mpad], the peers'
   * blocks opened through CUDA IPC, one copy stream per peer */
  double *d_gather = nullptr;
  size_t gather_elems_per_rank = 0;
  std::vector<double *> peer_gather;
  std::vector<cudaStream_t> push_stream;
  int xchg_self = -1;

  int fail(int code, const std::string &msg) {
    err = msg;
    return code;
  }

  /* per-member N2O / halocarbon parameters (the GAS build): values given per member, by GP_* id;
   * scalars stay in UC_N2O / TN2O0 / halo_* above */
  std::vector<double> gvec[GP_COUNT];
  double *d_GP = nullptr, *d_GF = nullptr, *d_GF_snap = nullptr, *d_scen_gas = nullptr;
  bool gas_per_member() const {
    if (!pvec[PI_N0].empty()) return true;
    for (const std::vector<double> &v : gvec)
      if (!v.empty()) return true;
    return false;
  }
  double gas_scalar(int gp) const {
    if (gp == GP_UC_N2O) return UC_N2O;
    if (gp == GP_TN2O0) return TN2O0;
    const int g = (gp - GP_HALO0) / 5, f = (gp - GP_HALO0) % 5;
    const double *tab[5] = {halo_tau, halo_rho, halo_delta, halo_H0, halo_mm};
    return tab[f][g];
  }
  /* "UC_N2O", "TN2O0", "<gas>.tau|rho|delta|H0|molarMass" -> GP_* id, or -1 */
  static int find_gas_param(const char *name) {
    if (!strcmp(name, "UC_N2O")) return GP_UC_N2O;
    if (!strcmp(name, "TN2O0")) return GP_TN2O0;
    const char *dot = strchr(name, '.');
    if (!dot) return -1;
    const std::string gas(name, dot - name), field(dot + 1);
    static const char *const fields[5] = {"tau", "rho", "delta", "H0", "molarMass"};
    for (int g = 0; g < HX_NHALO; ++g)
      if (gas == hx::kHaloNames[g])
        for (int f = 0; f < 5; ++f)
          if (field == fields[f]) return GP_HALO0 + 5 * g + f;
    return -1;
  }
  double gas_value(int gp, int member) const { return gvec[gp].empty() ? gas_scalar(gp) : gvec[gp][member]; }
  int upload_gas() {
    for (int gp = 0; gp < GP_COUNT; ++gp) {
      k_fill_field<<<(Mpad + 255) / 256, 256, 0, stream>>>(d_GP, gp, GP_COUNT, gas_scalar(gp), Mpad);
      CUDA_TRY(cudaGetLastError());
      if (gvec[gp].empty()) continue;
      int rc = ensure_pinned((size_t)M * sizeof(double));
      if (rc) return rc;
      rc = ensure_stage((size_t)M * sizeof(double));
      if (rc) return rc;
      CUDA_TRY(cudaStreamSynchronize(stream));
      memcpy(h_pinned, gvec[gp].data(), (size_t)M * sizeof(double));
      CUDA_TRY(cudaMemcpyAsync(d_stage, h_pinned, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, stream));
      k_scatter_field<<<(M + 255) / 256, 256, 0, stream>>>(d_GP, gp, GP_COUNT, d_stage, d_dev_of_api, M);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return HX_OK;
  }
  /* one member's halocarbon series with its own parameters (the derived outputs of the GAS
   * build): concentration and adjusted forcing per row, halocarbon_component.cpp:181-229 */
  void halo_series_member(int s, int g, int member, std::vector<double> &conc, std::vector<double> &rf) const {
    const double tau = gas_value(GP_HALO0 + 5 * g + 0, member), rho = gas_value(GP_HALO0 + 5 * g + 1, member),
                 delta = gas_value(GP_HALO0 + 5 * g + 2, member), mm = gas_value(GP_HALO0 + 5 * g + 4, member);
    double Ha = gas_value(GP_HALO0 + 5 * g + 3, member);
    const double *E = raw[s].data() + (size_t)(RAW_HALO0 + g) * nrow;
    conc.assign(nrow, 0.0);
    rf.assign(nrow, 0.0);
    const double expfac = std::exp(-(1 / tau));
    for (int r = 1; r < nrow; ++r) {
      const double emissMol = E[r] / mm * 1.0;
      const double concDeltaEmiss = emissMol / (0.1 * 1.8);
      Ha = Ha * expfac + concDeltaEmiss * tau * (1.0 - expfac);
      conc[r] = Ha;
      const double rf_unadjusted = rho * Ha;
      rf[r] = rf_unadjusted + delta * rf_unadjusted;
    }
  }

  /* biomes: hx_set_biomes */
  int n_biomes = 1;
  std::vector<std::string> biome_names;
  double bscalar[HX_MAX_BIOMES][BP_COUNT];
  std::vector<double> bvec[HX_MAX_BIOMES][BP_COUNT];
  double *d_BP = nullptr, *d_BF = nullptr, *d_BF_snap = nullptr;
  /* "<biome>.<name>" -> (biome, BP_* field); false if it is not one */
  bool find_biome_param(const char *name, int &ib, int &f) const {
    const char *dot = strchr(name, '.');
    if (!dot || n_biomes <= 1) return false;
    const std::string biome(name, dot - name);
    for (ib = 0; ib < n_biomes; ++ib)
      if (biome_names[ib] == biome) break;
    if (ib == n_biomes) return false;
    for (f = 0; f < BP_COUNT; ++f)
      if (!strcmp(hx::kBiomeParams[f].name, dot + 1)) return true;
    return false;
  }
  /* the global land inputs that per-biome values replace (simpleNbox.cpp:201-227) */
  static bool is_biome_replaced(int pi) {
    static const int k[] = {PI_VEG_C0, PI_DET_C0, PI_SOIL_C0, PI_PERMAFROST_C0, PI_NPP_FLUX0, PI_BETA,
                            PI_Q10, PI_WARMINGFACTOR, PI_F_NPPV, PI_F_NPPD, PI_F_LITTERD,
                            PI_RH_CH4_FRAC, PI_PF_MU, PI_PF_SIGMA, PI_FPF_STATIC};
    for (int v : k)
      if (v == pi) return true;
    return false;
  }
  int upload_biomes() {
    const int nf = n_biomes * BP_COUNT;
    for (int ib = 0; ib < n_biomes; ++ib)
      for (int f = 0; f < BP_COUNT; ++f) {
        const int idx = ib * BP_COUNT + f;
        k_fill_field<<<(Mpad + 255) / 256, 256, 0, stream>>>(d_BP, idx, nf, bscalar[ib][f], Mpad);
        CUDA_TRY(cudaGetLastError());
        if (bvec[ib][f].empty()) continue;
        int rc = ensure_pinned((size_t)M * sizeof(double));
        if (rc) return rc;
        rc = ensure_stage((size_t)M * sizeof(double));
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(stream));
        memcpy(h_pinned, bvec[ib][f].data(), (size_t)M * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(d_stage, h_pinned, (size_t)M * sizeof(double),
                                 cudaMemcpyHostToDevice, stream));
        k_scatter_field<<<(M + 255) / 256, 256, 0, stream>>>(d_BP, idx, nf, d_stage, d_dev_of_api, M);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(stream));
      }
    return HX_OK;
  }

  int find_param(const char *name) const {
    for (int i = 0; i < PI_COUNT; ++i)
      if (!strcmp(hx::kParams[i].name, name)) return i;
    return -1;
  }
  static int find_raw(const char *name) {
    for (int i = 0; i < RAW_HALO0; ++i)
      if (!strcmp(hx::kRawNames[i], name)) return i;
    for (int g = 0; g < HX_NHALO; ++g) {
      std::string s = std::string(hx::kHaloNames[g]) + "_emissions";
      if (s == name) return RAW_HALO0 + g;
    }
    return -1;
  }
  static int find_constraint(const char *name) {
    static const char *const base[CN_HALO0] = {"CO2_constrain", "NBP_constrain", "CH4_constrain",
                                               "N2O_constrain", "RF_tot_constrain",
                                               "tas_constrain"};
    for (int i = 0; i < CN_HALO0; ++i)
      if (!strcmp(base[i], name)) return i;
    for (int g = 0; g < HX_NHALO; ++g) {
      std::string s = std::string(hx::kHaloNames[g]) + "_constrain";
      if (s == name) return CN_HALO0 + g;
    }
    return -1;
  }
  const double *con(int s, int series) const { return cons[s].data() + (size_t)series * nrow; }
  /* tseries::get() of a series that allows interpolation (tas_constrain, RF_tot_constrain:
   * temperature_component.cpp:112, forcing_component.cpp:112): linear between entries;
   * `flat_below` extends the first entry downwards (RF_tot is applied for every year up to the
   * last entry, forcing_component.cpp:498; tas only between first and last, :510-512) */
  std::vector<double> densify(int s, int series, bool flat_below) const {
    const double *c = con(s, series);
    std::vector<double> out(nrow, NAN);
    int first = -1, last = -1;
    for (int r = 0; r < nrow; ++r)
      if (c[r] == c[r]) { if (first < 0) first = r; last = r; }
    if (first < 0) return out;
    int lo = first;
    for (int r = flat_below ? 0 : first; r <= last; ++r) {
      if (r < first) { out[r] = c[first]; continue; }
      if (c[r] == c[r]) { out[r] = c[r]; lo = r; continue; }
      int hi = r;
      while (!(c[hi] == c[hi])) ++hi;
      const double x0 = cfg.start_year + lo, x1 = cfg.start_year + hi, t = cfg.start_year + r;
      out[r] = c[lo] + (t - x0) * (c[hi] - c[lo]) / (x1 - x0);
    }
    return out;
  }

  /* output id of a name: OUT_* or, with biomes, "<biome>.<name>" -> OUT_COUNT + ... */
  int find_out(const char *name) const {
    for (int i = 0; i < OUT_COUNT; ++i)
      if (!strcmp(hx::kOutNames[i], name)) return i;
    const char *dot = name ? strchr(name, '.') : nullptr;
    if (dot && n_biomes > 1) {
      const std::string biome(name, dot - name);
      for (int ib = 0; ib < n_biomes; ++ib)
        if (biome_names[ib] == biome)
          for (int k = 0; k < BO_COUNT; ++k)
            if (!strcmp(hx::kBiomeOutNames[k], dot + 1)) return OUT_COUNT + ib * BO_COUNT + k;
    }
    return -1;
  }

  int ensure_stage(size_t bytes) {
    if (bytes <= stage_bytes) return HX_OK;
    if (d_stage) cudaFree(d_stage);
    d_stage = nullptr;
    stage_bytes = 0;
    CUDA_TRY(cudaMalloc(&d_stage, bytes));
    stage_bytes = bytes;
    return HX_OK;
  }
  /* field f of a CTA-tiled device array for every member, API order, on the host */
  int gather_field_host(const double *A, int f, int nfields, double *out) {
    int rc = ensure_stage((size_t)M * sizeof(double));
    if (rc) return rc;
    k_gather_field<<<(M + 255) / 256, 256, 0, stream>>>(d_stage, A, f, nfields, d_dev_of_api, M);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, d_stage, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return HX_OK;
  }
  /* parameter pi of every member, API order */
  int param_values(int pi, double *out) {
    if (pvec_on_device_only[pi]) return gather_field_host(d_P, pi, PD_COUNT, out);
    for (int i = 0; i < M; ++i) out[i] = pvec[pi].empty() ? pscalar[pi] : pvec[pi][i];
    return HX_OK;
  }
  int start_row(int id, std::vector<double> &row);
  int fetch_rows(int slot, const std::vector<int32_t> &yidx, int n_dates, double *out);
  int ensure_pinned(size_t bytes) {
    if (bytes <= pinned_bytes) return HX_OK;
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    pinned_bytes = 0;
    CUDA_TRY(cudaMallocHost(&h_pinned, bytes));
    pinned_bytes = bytes;
    return HX_OK;
  }

  /* ---- member-independent gas series on the host (E-5) ----
   * N2OComponent::run (n2o_component.cpp:150-191): tau = TN2O0 (N2O/N0)^-0.05,
   *   dN2O = (E + E_nat)/UC_N2O - N2O/tau
   * HalocarbonComponent::run (halocarbon_component.cpp:181-229): exponential decay + emissions,
   *   RF = rho Ha (1 + delta) */
  void gas_series(int s, std::vector<double> &n2o, std::vector<double> &halo_rf,
                  std::vector<double> *halo_conc = nullptr) const {
    if (halo_conc) halo_conc->assign((size_t)nrow * HX_NHALO, 0.0);
    const double *R = raw[s].data();
    auto rawv = [&](int series, int r) { return R[(size_t)series * nrow + r]; };
    double N0 = pscalar[PI_N0];
    const double *cn2o = con(s, CN_N2O);
    if (cn2o[0] == cn2o[0]) N0 = cn2o[0]; /* n2o_component.cpp:141-146 */
    n2o.assign(nrow, 0.0);
    halo_rf.assign((size_t)nrow * HX_NHALO, 0.0);
    n2o[0] = N0;
    for (int r = 1; r < nrow; ++r) {
      if (cn2o[r] == cn2o[r]) { /* concentration-forced year, :157-158 */
        n2o[r] = cn2o[r];
        continue;
      }
      const double previous_n2o = n2o[r - 1];
      const double tau = TN2O0 * (std::pow(previous_n2o / N0, -0.05));
      const double current_n2oem = rawv(RAW_N2O_E, r) + rawv(RAW_N2O_NAT, r);
      const double dN2O = current_n2oem / UC_N2O - previous_n2o / tau;
      n2o[r] = previous_n2o + dN2O;
    }
    for (int g = 0; g < HX_NHALO; ++g) {
      double Ha = halo_H0[g];
      const double tau = halo_tau[g];
      const double *cha = con(s, CN_HALO0 + g);
      for (int r = 1; r < nrow; ++r) {
        const double timestep = 1.0;
        const double alpha = 1 / tau;
        const double emissMol = rawv(RAW_HALO0 + g, r) / halo_mm[g] * timestep;
        const double concDeltaEmiss = emissMol / (0.1 * 1.8);
        const double expfac = std::exp(-alpha);
        Ha = Ha * expfac + concDeltaEmiss * tau * (1.0 - expfac);
        if (cha[r] == cha[r]) Ha = cha[r]; /* halocarbon_component.cpp:189-192 */
        if (halo_conc) (*halo_conc)[(size_t)r * HX_NHALO + g] = Ha;
        const double rf_unadjusted = halo_rho[g] * Ha;
        halo_rf[(size_t)r * HX_NHALO + g] = rf_unadjusted + halo_delta[g] * rf_unadjusted;
      }
    }
  }

  /* Outputs that need no kernel: the forcing of every agent other than CO2 / CH4 / N2O, the
   * halocarbon concentrations and forcings.  They are functions of the scenario series, of
   * per-member parameters and (two of them) of recorded outputs, evaluated at fetch time with the
   * expressions of ForcingComponent::run (forcing_component.cpp:410-489), relative to the base
   * year like every forcing the reference reports (:507-524); before the base year they are 0.
   *   RF_BC RF_OC RF_SO2 RF_NH3 RF_aci RF_vol RF_albedo RF_misc RF_O3_trop RF_H2O_strat
   *   RF_<gas> (absolute, as the halocarbon component reports it), Fadj<gas> (relative),
   *   <gas>_concentration
   * Returns 1 if `name` is one of them (rc holds the result code), 0 otherwise. */
  int fetch_derived(const char *name, const double *dates, int n_dates, double *out, int &rc);
  int fetch_functions(const char *name, const double *dates, int n_dates, double *out, int &rc);

  void free_device() {
    void *ptrs[] = {d_GP, d_GF, d_GF_snap, d_scen_gas, d_BP, d_BF, d_BF_snap, d_P, d_S, d_S_snap, d_ker, d_conv, d_sst, d_tland, d_out, d_X, d_scen, d_stage,
                    d_block_scen, d_status, d_status_snap, d_status_post, d_fail_year, d_spinup_steps, d_yidx,
                    d_counters, d_dev_of_api, d_api_of_dev, d_sched, d_T, d_TO, d_TK, d_TOK, d_REC, d_YCNT, d_trk_fail};
    for (void *p : ptrs)
      if (p) cudaFree(p);
    d_P = d_S = d_S_snap = d_D = d_ker = d_conv = d_sst = d_tland = d_out = d_X = d_scen = d_stage = nullptr;
    d_block_scen = d_status = d_status_snap = d_status_post = d_fail_year = d_spinup_steps = d_yidx = nullptr;
    d_counters = nullptr;
    d_sched = nullptr;
    d_T = d_TO = d_REC = nullptr;
    d_YCNT = nullptr;
    d_TK = d_TOK = nullptr;
    d_trk_fail = nullptr;
    d_BP = d_BF = d_BF_snap = nullptr;
    d_GP = d_GF = d_GF_snap = d_scen_gas = nullptr;
    d_dev_of_api = d_api_of_dev = nullptr;
    stage_bytes = 0;
    yidx_cap = 0;
  }

  int upload_param(int pi) {
    if (pvec_on_device_only[pi]) return HX_OK; /* already resident (hx_set_param_device) */
    k_fill_field<<<(Mpad + 255) / 256, 256, 0, stream>>>(d_P, pi, PD_COUNT, pscalar[pi], Mpad);
    CUDA_TRY(cudaGetLastError());
    if (!pvec[pi].empty()) {
      int rc = ensure_pinned((size_t)M * sizeof(double));
      if (rc) return rc;
      rc = ensure_stage((size_t)M * sizeof(double));
      if (rc) return rc;
      /* the pinned buffer may still feed an earlier async copy */
      CUDA_TRY(cudaStreamSynchronize(stream));
      memcpy(h_pinned, pvec[pi].data(), (size_t)M * sizeof(double));
      CUDA_TRY(cudaMemcpyAsync(d_stage, h_pinned, (size_t)M * sizeof(double),
                               cudaMemcpyHostToDevice, stream));
      k_scatter_field<<<(M + 255) / 256, 256, 0, stream>>>(d_P, pi, PD_COUNT, d_stage, d_dev_of_api, M);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return HX_OK;
  }

  /* device scenario tables [scen][row][SC_STRIDE]: raw series, host gas series, constraints */
  void build_tables(std::vector<double> &tab, bool &any_constraint, bool &any_nbp) const {
    tab.assign((size_t)nscen * nrow * SC_STRIDE, 0.0);
    any_constraint = false;
    any_nbp = false;
    for (int s = 0; s < nscen; ++s) {
      std::vector<double> n2o, hrf;
      gas_series(s, n2o, hrf);
      const double *R = raw[s].data();
      const std::vector<double> rftot = densify(s, CN_RFTOT, true), tas = densify(s, CN_TAS, false);
      const double *cco2 = con(s, CN_CO2), *cch4 = con(s, CN_CH4), *cnbp = con(s, CN_NBP);
      for (int r = 0; r < nrow; ++r) {
        double *row = tab.data() + ((size_t)s * nrow + r) * SC_STRIDE;
        for (int c = 0; c <= SC_MISC; ++c) row[c] = R[(size_t)c * nrow + r]; /* RAW_x == SC_x up to MISC */
        row[SC_N2O] = n2o[r];
        row[SC_SQRT_N2O] = std::sqrt(n2o[r]);
        {
          const double aci_beta = 2.279759, s_BCOC = 111.05064063,
                       s_SO2 = (260.34644166 * 1000) * (32.065 / 64.066);
          const double E_BC = row[SC_BC], E_OC = row[SC_OC], E_SO2 = row[SC_SO2];
          row[SC_ACI] = -1 * aci_beta * std::log(1 + (E_SO2 / s_SO2) + ((E_BC + E_OC) / s_BCOC));
        }
        for (int g = 0; g < HX_NHALO; ++g) row[SC_HALO0 + g] = hrf[(size_t)r * HX_NHALO + g];
        row[SC_C_CO2] = cco2[r]; row[SC_C_CH4] = cch4[r];
        row[SC_C_RFTOT] = rftot[r]; row[SC_C_TAS] = tas[r]; row[SC_C_NBP] = cnbp[r];
        for (int c = SC_C_CO2; c <= SC_C_NBP; ++c)
          if (row[c] == row[c]) any_constraint = true;
        if (row[SC_C_NBP] == row[SC_C_NBP]) any_nbp = true;
      }
    }
  }
  int upload_tables() {
    std::vector<double> tab;
    bool any = false, nbp = false;
    build_tables(tab, any, nbp);
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(d_scen, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    tables_constrained = any;
    tables_nbp = nbp;
    tables_dirty = false;
    return HX_OK;
  }

  bool spinup_shared() const {
    if (n_biomes > 1) return false; /* the broadcast covers the member-level state only */
    for (int pi : hx::kSpinupParams)
      if (!pvec[pi].empty() || pvec_on_device_only[pi]) return false;
    return true;
  }

  int first_active = 0;

  /* rows ra+1 .. rb of the run.  With carbon tracking on the run kernel only records each
   * stash's flux scalars, one 16-year slab per launch, and the replay kernel folds them into
   * the source maps before the next slab overwrites the record. */
  cudaStream_t track_stream = nullptr;
  cudaEvent_t ev_rec_full[2] = {nullptr, nullptr}, ev_rec_free[2] = {nullptr, nullptr};
  size_t rec_elems = 0, ycnt_bytes = 0; /* size of ONE slab's record; a buffer holds rec_group of them */
  int rec_group = 1;                    /* slabs per tracked launch of the run kernel */

  /* The record is double buffered and the replay runs on a stream of its own, so the run
   * kernel of one group of slabs overlaps the replay of the group before (the replay's CTAs
   * fill the SMs that the run kernel's last wave leaves idle, and vice versa).  Replays stay in
   * order on their stream; a record buffer is rewritten only after its replays have finished.
   * A launch of the persistent run kernel covers rec_group slabs (as many as the device memory
   * left for the record allows, up to 4): the claim scheduler keeps every SM busy across them,
   * where one launch per slab ended in a partial wave each time. */
  cudaError_t launch_rows(int ra, int rb) {
    if (!d_T) return hx::launch_run(d, C, ra, rb, stream);
    if (!track_stream) {
      cudaError_t e = cudaStreamCreateWithFlags(&track_stream, cudaStreamNonBlocking);
      for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ev_rec_full[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_rec_free[i], cudaEventDisableTiming);
      }
      if (e != cudaSuccess) return e;
    }
    const int slab = hx::track_slab_years();
    bool used[2] = {false, false};
    for (int i = 0; ra < rb; ++i) {
      const int re = std::min(rb, ra + slab * rec_group);
      const int b = i & 1;
      HxDev db = d;
      db.REC = d_REC + (size_t)b * rec_group * rec_elems;
      db.YCNT = d_YCNT + (size_t)b * rec_group * ycnt_bytes;
      db.rec_slab_stride = rec_elems;
      db.ycnt_slab_stride = ycnt_bytes;
      cudaError_t e = cudaSuccess;
      if (used[b]) e = cudaStreamWaitEvent(stream, ev_rec_free[b], 0);
      if (e == cudaSuccess) e = hx::launch_run(db, C, ra, re, stream);
      if (e == cudaSuccess && cfg.start_year + re >= C.tracking_date) {
        e = cudaEventRecord(ev_rec_full[b], stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(track_stream, ev_rec_full[b], 0);
        /* one replay per slab of the group, in order */
        for (int g = 0, sa = ra; sa < re && e == cudaSuccess; ++g, sa += slab) {
          const int se = std::min(re, sa + slab);
          if (cfg.start_year + se < C.tracking_date) continue;
          HxDev dg = db;
          dg.REC = db.REC + (size_t)g * rec_elems;
          dg.YCNT = db.YCNT + (size_t)g * ycnt_bytes;
          e = hx::launch_track(dg, C, sa, se, track_stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev_rec_free[b], track_stream);
        used[b] = true;
      }
      if (e != cudaSuccess) return e;
      ra = re;
    }
    /* whatever follows on the engine's stream sees the maps of every replay; the replays
     * report a bad mix through a word of their own (the run kernel of the next slab reads and
     * writes the status words meanwhile), folded into the status here, in stream order */
    for (int b = 0; b < 2; ++b)
      if (used[b]) {
        cudaError_t e = cudaStreamWaitEvent(stream, ev_rec_free[b], 0);
        if (e != cudaSuccess) return e;
      }
    if (used[0] || used[1]) return hx::launch_track_merge(d, stream);
    return cudaSuccess;
  }

  /* hx_run_stream, untracked: see there.  slab_done lives in mapped pinned memory; the last
   * tile to finish slab s sets slab_done[s] (tiles are counted in device memory; the store
   * follows a system-scope fence), this thread polls it and queues the slab's device-to-host
   * copies on the copy stream. */
  unsigned *h_slab_done = nullptr, *d_slab_done = nullptr;
  int slab_done_cap = 0;
  int run_streamed(int r0, int r1, int n_vars, const int *slot, double *const *outs) {
    const int nslab = (r1 - r0 + HX_SLAB_YEARS - 1) / HX_SLAB_YEARS;
    if (nslab > slab_done_cap) {
      if (h_slab_done) cudaFreeHost(h_slab_done);
      h_slab_done = nullptr;
      slab_done_cap = 0;
      CUDA_TRY(cudaHostAlloc((void **)&h_slab_done, (size_t)nslab * sizeof(unsigned), cudaHostAllocMapped));
      CUDA_TRY(cudaHostGetDevicePointer((void **)&d_slab_done, h_slab_done, 0));
      slab_done_cap = nslab;
    }
    memset(h_slab_done, 0, (size_t)nslab * sizeof(unsigned));
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, HX_NCOUNTERS * sizeof(unsigned long long), stream));
    CUDA_TRY(cudaEventRecord(ev0, stream));
    HxDev ds = d;
    ds.slab_done = d_slab_done;
    CUDA_TRY(hx::launch_run(ds, C, r0, r1, stream));
    CUDA_TRY(cudaEventRecord(ev1, stream));
    volatile unsigned *done = h_slab_done; /* done[s] becomes 1: raised by slab s's last tile */
    bool kernel_over = false;
    for (int s = 0; s < nslab; ++s) {
      unsigned spins = 0;
      while (done[s] == 0u && !kernel_over) {
        if ((++spins & 0xfffu) == 0) {
          /* the kernel may have stopped without finishing (a launch or device error) */
          const cudaError_t q = cudaStreamQuery(stream);
          if (q == cudaSuccess) kernel_over = true;
          else if (q != cudaErrorNotReady) return fail(HX_ERR_CUDA, std::string("run kernel: ") + cudaGetErrorString(q));
        }
      }
      if (done[s] == 0u)
        return fail(HX_ERR_CUDA, "hx_run_stream: the run kernel ended without completing every slab");
      const int ra = r0 + s * HX_SLAB_YEARS, rb = std::min(r1, ra + HX_SLAB_YEARS);
      for (int v = 0; v < n_vars; ++v) {
        const double *src = d_out + ((size_t)slot[v] * (nrow - 1) + ra) * Mpad;
        double *dst = outs[v] + (size_t)(ra - r0) * M;
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)M * sizeof(double), src, (size_t)Mpad * sizeof(double),
                                   (size_t)M * sizeof(double), (size_t)(rb - ra),
                                   cudaMemcpyDeviceToHost, copy_stream));
      }
    }
    CUDA_TRY(cudaEventRecord(ev_copy, copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(stream, ev_copy, 0));
    CUDA_TRY(cudaStreamSynchronize(copy_stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return HX_OK;
  }

  /* hx_run_exchange: like run_streamed, but a finished slab's rows go to every peer's gather
   * block (and to this rank's own) instead of to the host: device-to-device copies on the copy
   * engines, over NVLink for the peers, one stream per destination so that they spread over the
   * engines, while the kernel computes the later slabs. */
  int run_pushed(int r0, int r1) {
    const int nslab = (r1 - r0 + HX_SLAB_YEARS - 1) / HX_SLAB_YEARS;
    if (nslab > slab_done_cap) {
      if (h_slab_done) cudaFreeHost(h_slab_done);
      h_slab_done = nullptr;
      slab_done_cap = 0;
      CUDA_TRY(cudaHostAlloc((void **)&h_slab_done, (size_t)nslab * sizeof(unsigned), cudaHostAllocMapped));
      CUDA_TRY(cudaHostGetDevicePointer((void **)&d_slab_done, h_slab_done, 0));
      slab_done_cap = nslab;
    }
    memset(h_slab_done, 0, (size_t)nslab * sizeof(unsigned));
    CUDA_TRY(cudaMemsetAsync(d_counters, 0, HX_NCOUNTERS * sizeof(unsigned long long), stream));
    CUDA_TRY(cudaEventRecord(ev0, stream));
    HxDev ds = d;
    ds.slab_done = d_slab_done;
    CUDA_TRY(hx::launch_run(ds, C, r0, r1, stream));
    CUDA_TRY(cudaEventRecord(ev1, stream));
    volatile unsigned *done = h_slab_done;
    const int n = (int)peer_gather.size(), nsel = (int)out_sel.size();
    const size_t ny = (size_t)(nrow - 1);
    bool kernel_over = false;
    for (int s = 0; s < nslab; ++s) {
      unsigned spins = 0;
      while (done[s] == 0u && !kernel_over) {
        if ((++spins & 0xfffu) == 0) {
          const cudaError_t q = cudaStreamQuery(stream);
          if (q == cudaSuccess) kernel_over = true;
          else if (q != cudaErrorNotReady) return fail(HX_ERR_CUDA, std::string("run kernel: ") + cudaGetErrorString(q));
        }
      }
      if (done[s] == 0u)
        return fail(HX_ERR_CUDA, "hx_run_exchange: the run kernel ended without completing every slab");
      const int ra = r0 + s * HX_SLAB_YEARS, rb = std::min(r1, ra + HX_SLAB_YEARS);
      for (int k = 0; k < n; ++k) {
        const int p = (xchg_self + 1 + k) % n; /* every rank starts with a different peer */
        for (int v = 0; v < nsel; ++v) {
          const double *src = d_out + ((size_t)v * ny + ra) * Mpad;
          double *dst = peer_gather[p] + (size_t)xchg_self * gather_elems_per_rank + ((size_t)v * ny + ra) * Mpad;
          CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)(rb - ra) * Mpad * sizeof(double), cudaMemcpyDefault,
                                   push_stream[p]));
        }
      }
    }
    for (cudaStream_t ps : push_stream) CUDA_TRY(cudaStreamSynchronize(ps));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return HX_OK;
  }

  int run_setup_and_spinup() {
    const bool tables_changed = tables_dirty; /* a CH4 constraint in the first row sets the initial CH4 */
    if (tables_dirty) {
      int rc = upload_tables();
      if (rc) return rc;
    }
    /* a land-ocean warming ratio is served by the constraint builds of the run kernel */
    bool lo_active = pscalar[PI_LO_RATIO] != 0.0 || pvec_on_device_only[PI_LO_RATIO];
    for (double v : pvec[PI_LO_RATIO]) lo_active = lo_active || v != 0.0;
    d.constrained = tables_nbp ? 2 : (tables_constrained || lo_active) ? 1 : 0;
    if (!spinup_dirty && !tables_changed) {
      /* nothing the spin-up depends on changed: the post-spin-up snapshot still stands */
      CUDA_TRY(cudaMemcpyAsync(d_S, d_S_snap, (size_t)SI_COUNT * Mpad * sizeof(double),
                               cudaMemcpyDeviceToDevice, stream));
      if (d_BF)
        CUDA_TRY(cudaMemcpyAsync(d_BF, d_BF_snap, (size_t)n_biomes * BF_COUNT * Mpad * sizeof(double),
                                 cudaMemcpyDeviceToDevice, stream));
      if (d_GF)
        CUDA_TRY(cudaMemcpyAsync(d_GF, d_GF_snap, (size_t)GF_COUNT * Mpad * sizeof(double),
                                 cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(cudaMemcpyAsync(d_status, d_status_post, (size_t)Mpad * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(hx::launch_setup(d, C, stream, 2)); /* DOECLIM matrices, lag kernel, derived constants */
      if (d_T) CUDA_TRY(hx::launch_track_init(d, stream));
      if (d_trk_fail) CUDA_TRY(cudaMemsetAsync(d_trk_fail, 0, (size_t)Mpad * sizeof(int32_t), stream));
      cur_row = 0;
      params_dirty = false;
      dirty_from_row = 0x7fffffff;
      return HX_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(d_status, d_status_snap, (size_t)Mpad * sizeof(int32_t),
                             cudaMemcpyDeviceToDevice, stream));
    if (spinup_shared() && M > 1) {
      /* E-7: nothing that shapes the spin-up or the alkalinity equilibration varies across
       * members, so one member's spin-up serves the ensemble: run it on one thread and
       * broadcast the state rows it touches (the first SI_SPINUP_ROWS).  That single thread is
       * pure latency (2.2 ms), so the members' DOECLIM set-up (1.0 ms, reads and writes nothing
       * the spin-up touches) runs beside it on a second stream. */
      if (!setup_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&setup_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_setup_a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_setup_b, cudaEventDisableTiming));
      }
      CUDA_TRY(hx::launch_setup(d, C, stream, 1));
      CUDA_TRY(cudaEventRecord(ev_setup_a, stream));
      CUDA_TRY(cudaStreamWaitEvent(setup_stream, ev_setup_a, 0));
      CUDA_TRY(hx::launch_setup(d, C, setup_stream, 2));
      CUDA_TRY(cudaEventRecord(ev_setup_b, setup_stream));
      CUDA_TRY(hx::launch_spinup_one(d, C, first_active, stream));
      k_broadcast_state<<<(Mpad + 255) / 256, 256, 0, stream>>>(
          d_S, d_spinup_steps, d_status, d_fail_year, first_active, SI_SPINUP_ROWS, (size_t)Mpad);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_setup_b, 0));
    } else {
      CUDA_TRY(hx::launch_setup(d, C, stream));
      CUDA_TRY(hx::launch_spinup(d, C, stream));
    }
    if (d_T) CUDA_TRY(hx::launch_track_init(d, stream));
    if (d_trk_fail) CUDA_TRY(cudaMemsetAsync(d_trk_fail, 0, (size_t)Mpad * sizeof(int32_t), stream));
    CUDA_TRY(cudaMemcpyAsync(d_S_snap, d_S, (size_t)SI_COUNT * Mpad * sizeof(double),
                             cudaMemcpyDeviceToDevice, stream));
    if (d_BF)
      CUDA_TRY(cudaMemcpyAsync(d_BF_snap, d_BF, (size_t)n_biomes * BF_COUNT * Mpad * sizeof(double),
                               cudaMemcpyDeviceToDevice, stream));
    if (d_GF)
      CUDA_TRY(cudaMemcpyAsync(d_GF_snap, d_GF, (size_t)GF_COUNT * Mpad * sizeof(double),
                               cudaMemcpyDeviceToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_status_post, d_status, (size_t)Mpad * sizeof(int32_t),
                             cudaMemcpyDeviceToDevice, stream));
    cur_row = 0;
    params_dirty = false;
    dirty_from_row = 0x7fffffff;
    spinup_dirty = false;
    return HX_OK;
  }
};

} // namespace

struct hx_engine : Engine {};

int Engine::fetch_derived(const char *name, const double *dates, int n_dates, double *out, int &rc) {
  const std::string nm(name);
  enum Kind { K_NONE, K_BC, K_OC, K_SO2, K_NH3, K_ACI, K_VOL, K_ALBEDO, K_MISC, K_O3, K_H2O,
              K_HALO_RF, K_HALO_ADJ, K_HALO_CONC };
  Kind kind = K_NONE;
  int gas = -1;
  static const struct { const char *n; Kind k; } simple[] = {
      {"RF_BC", K_BC}, {"RF_OC", K_OC}, {"RF_SO2", K_SO2}, {"RF_NH3", K_NH3}, {"RF_aci", K_ACI},
      {"RF_vol", K_VOL}, {"RF_albedo", K_ALBEDO}, {"RF_misc", K_MISC}, {"RF_O3_trop", K_O3},
      {"RF_H2O_strat", K_H2O}};
  for (const auto &e : simple)
    if (nm == e.n) kind = e.k;
  for (int g = 0; g < HX_NHALO && kind == K_NONE; ++g) {
    const std::string gname(hx::kHaloNames[g]);
    if (nm == "RF_" + gname) { kind = K_HALO_RF; gas = g; }
    else if (nm == "Fadj" + gname) { kind = K_HALO_ADJ; gas = g; }
    else if (nm == gname + "_concentration") { kind = K_HALO_CONC; gas = g; }
  }
  if (kind == K_NONE) return 0;
  rc = HX_OK;
  const int base = C.baseyear - cfg.start_year; /* base-year row */
  std::vector<int> rows(n_dates);
  for (int k = 0; k < n_dates; ++k) {
    rows[k] = (int)dates[k] - cfg.start_year;
    if (rows[k] < 1 || rows[k] > cur_row) {
      rc = fail(HX_ERR_ARG, "date outside (start_year, current date]");
      return 1;
    }
  }
  /* per-member parameters and statuses */
  auto param = [&](int pi, std::vector<double> &v) {
    v.resize(M);
    return hx_get_param(static_cast<hx_engine *>(this), hx::kParams[pi].name, v.data(), M);
  };
  std::vector<int32_t> st(M), fy(M);
  rc = hx_member_status(static_cast<hx_engine *>(this), st.data(), fy.data(), M);
  if (rc) return 1;
  /* recorded outputs some agents are functions of (at the requested dates and the base year) */
  std::vector<double> rec, rec_base;
  auto need_output = [&](const char *var) {
    std::vector<double> d2(dates, dates + n_dates);
    rec.resize((size_t)M * n_dates);
    rec_base.assign(M, 0.0);
    int r2 = hx_fetch(static_cast<hx_engine *>(this), var, d2.data(), n_dates, rec.data());
    if (r2 == HX_OK && base >= 1 && base <= cur_row) {
      const double by = C.baseyear;
      r2 = hx_fetch(static_cast<hx_engine *>(this), var, &by, 1, rec_base.data());
    }
    return r2;
  };
  std::vector<double> aero, vol, rho, M0v;
  if (kind == K_BC || kind == K_OC || kind == K_SO2 || kind == K_NH3 || kind == K_ACI) {
    if ((rc = param(PI_AERO, aero))) return 1;
    const int pi = kind == K_BC ? PI_RHO_BC : kind == K_OC ? PI_RHO_OC : kind == K_SO2 ? PI_RHO_SO2 : PI_RHO_NH3;
    if (kind != K_ACI && (rc = param(pi, rho))) return 1;
  }
  if (kind == K_VOL && (rc = param(PI_VOL, vol))) return 1;
  if (kind == K_O3 && (rc = need_output("O3_concentration"))) return 1;
  if (kind == K_H2O) {
    if ((rc = need_output("CH4_concentration")) || (rc = param(PI_M0, M0v))) return 1;
  }
  /* per-scenario series */
  std::vector<std::vector<double>> hrf(nscen), hconc(nscen);
  if (kind == K_HALO_RF || kind == K_HALO_ADJ || kind == K_HALO_CONC) {
    std::vector<double> n2o;
    for (int s = 0; s < nscen; ++s) gas_series(s, n2o, hrf[s], &hconc[s]);
  }
  const double aci_beta = 2.279759, s_BCOC = 111.05064063,
               s_SO2 = (260.34644166 * 1000) * (32.065 / 64.066);
  const double nan = std::nan("");
  std::vector<double> m_conc, m_rf; /* GAS build: this member's own halocarbon series */
  for (int i = 0; i < M; ++i) {
    const int s = member_scen[i];
    const double *R = raw[s].data();
    auto rawv = [&](int series, int r) { return R[(size_t)series * nrow + r]; };
    if (d_GP && gas >= 0) halo_series_member(s, gas, i, m_conc, m_rf);
    /* absolute forcing of the agent in row r */
    auto absolute = [&](int r, int k) -> double {
      switch (kind) {
        case K_BC: return aero[i] * rho[i] * rawv(RAW_BC, r);
        case K_OC: return aero[i] * rho[i] * rawv(RAW_OC, r);
        case K_SO2: return aero[i] * rho[i] * rawv(RAW_SO2, r);
        case K_NH3: return aero[i] * rho[i] * rawv(RAW_NH3, r);
        case K_ACI:
          return aero[i] * (-1 * aci_beta *
                            std::log(1 + (rawv(RAW_SO2, r) / s_SO2) +
                                     ((rawv(RAW_BC, r) + rawv(RAW_OC, r)) / s_BCOC)));
        case K_VOL: return vol[i] * rawv(RAW_SV, r);
        case K_ALBEDO: return rawv(RAW_ALBEDO, r);
        case K_MISC: return rawv(RAW_MISC, r);
        case K_O3: return 0.042 * (k < 0 ? rec_base[i] : rec[(size_t)i * n_dates + k]);
        case K_H2O: {
          const double c0 = con(s, CN_CH4)[0];
          const double M0 = (c0 == c0) ? c0 : M0v[i]; /* ch4_component.cpp:141-146 */
          const double Ma = k < 0 ? rec_base[i] : rec[(size_t)i * n_dates + k];
          return 0.0485 * ((Ma - M0) / (1831 - M0));
        }
        case K_HALO_RF: case K_HALO_ADJ:
          return d_GP ? m_rf[r] : hrf[s][(size_t)r * HX_NHALO + gas];
        case K_HALO_CONC: return d_GP ? m_conc[r] : hconc[s][(size_t)r * HX_NHALO + gas];
        default: return nan;
      }
    };
    const bool relative = !(kind == K_HALO_RF || kind == K_HALO_CONC);
    for (int k = 0; k < n_dates; ++k) {
      const int r = rows[k];
      double v;
      if (st[i] > 0 && fy[i] <= cfg.start_year + r) v = nan; /* the member stopped before */
      else if (!relative) v = absolute(r, k);
      else if (r < base) v = 0.0; /* forcings are reported from the base year on */
      else v = absolute(r, k) - absolute(base, -1);
      out[(size_t)i * n_dates + k] = v;
    }
  }
  return 1;
}

/* The rest of the reference's outputstream variables that are plain functions of recorded
 * outputs, parameters and input series -- evaluated at fetch time with the reference's
 * expressions, no kernel involved (returns 0 when `name` is none of them):
 *   HL_sst, LL_sst   box temperature the year's chemistry ran at: last year's sst + 18 + deltaT
 *                    (oceanbox.cpp:97-99; ocean_component.cpp record_state)
 *   HL_DIC, LL_DIC   convertToDIC of the box's carbon (ocean_csys.cpp:403-408)
 *   DIC, pH, PCO2    area-weighted surface means, part_low x LL + part_high x HL
 *                    (ocean_component.cpp getData)
 *   ML_ocean_c       LL + HL carbon
 *   TAU_OH           OH lifetime (oh_component.cpp:137-174) from last year's CH4
 *   f_frozen         frozen permafrost fraction (simpleNbox-runtime.cpp:1012-1024) from last
 *                    year's land temperature; single biome only
 * Each needs the outputs it is a function of to be recorded (hx_select_outputs).  Pinned by
 *   HL_CO3, LL_CO3, CO3  carbonate of the year's last chemistry solve, from the recorded pCO2
 *                    and pH of that solve and the box temperature
 *   HL_OmegaCa, LL_OmegaCa, HL_OmegaAr, LL_OmegaAr   calcite / aragonite saturation from it
 * tests/golden/ref_outputs_more.npz.  (HL_ocean_uptake, LL_ocean_uptake, rh_det and rh_soil are
 * recorded by the run kernel when selected: hx_layout.h, "scratch rows X".) */
int Engine::fetch_functions(const char *name, const double *dates, int n_dates, double *out, int &rc) {
  enum Kind { F_NONE, F_SST_HL, F_SST_LL, F_DIC_HL, F_DIC_LL, F_DIC, F_PH, F_PCO2, F_ML, F_TAU_OH, F_FROZEN, F_CO3_HL, F_CO3_LL, F_CO3,
              F_OMEGACA_HL, F_OMEGACA_LL, F_OMEGAAR_HL, F_OMEGAAR_LL };
  static const struct { const char *n; Kind k; } tab[] = {
      {"HL_sst", F_SST_HL}, {"LL_sst", F_SST_LL}, {"HL_DIC", F_DIC_HL}, {"LL_DIC", F_DIC_LL},
      {"DIC", F_DIC}, {"pH", F_PH}, {"PCO2", F_PCO2}, {"ML_ocean_c", F_ML}, {"TAU_OH", F_TAU_OH},
      {"f_frozen", F_FROZEN}, {"HL_CO3", F_CO3_HL}, {"LL_CO3", F_CO3_LL}, {"CO3", F_CO3},
      {"HL_OmegaCa", F_OMEGACA_HL}, {"LL_OmegaCa", F_OMEGACA_LL}, {"HL_OmegaAr", F_OMEGAAR_HL},
      {"LL_OmegaAr", F_OMEGAAR_LL}};
  Kind kind = F_NONE;
  for (const auto &e : tab)
    if (!strcmp(name, e.n)) kind = e.k;
  int one_biome = -1; /* "<biome>.f_frozen": that biome's own fraction */
  if (kind == F_NONE && n_biomes > 1) {
    const char *dot = strchr(name, '.');
    if (dot && !strcmp(dot + 1, "f_frozen"))
      for (int ib = 0; ib < n_biomes; ++ib)
        if (biome_names[ib] == std::string(name, dot - name)) { kind = F_FROZEN; one_biome = ib; }
  }
  if (kind == F_NONE) return 0;
  rc = HX_OK;
  int rmax = 0;
  for (int k = 0; k < n_dates; ++k) {
    const int r = (int)dates[k] - cfg.start_year;
    if (r < 1 || r > cur_row) {
      rc = fail(HX_ERR_ARG, "date outside (start_year, current date]");
      return 1;
    }
    rmax = std::max(rmax, r);
  }
  hx_engine *self = static_cast<hx_engine *>(this);
  const size_t N = (size_t)M * n_dates;
  /* a recorded output at the requested dates, `back` years earlier (the start date included) */
  auto rec = [&](const char *var, int back, std::vector<double> &v) {
    std::vector<double> d2(n_dates);
    for (int k = 0; k < n_dates; ++k) d2[k] = dates[k] - back;
    v.resize(N);
    return hx_fetch(self, var, d2.data(), n_dates, v.data());
  };
  auto param = [&](int pi, std::vector<double> &v) {
    v.resize(M);
    return param_values(pi, v.data());
  };
  const double part_high = 0.15, part_low = 1 - part_high; /* ocean_component.hpp:89-90 */
  auto to_dic = [](double carbon, double volume) { /* ocean_csys.cpp:403-408, umol/kg */
    const double dic = ((carbon * 1e15) * (1.0 / 12.01) * (1.0 / 1027.0) * (1.0 / volume));
    return dic * 1e6;
  };
  std::vector<double> a, b;
  switch (kind) {
    case F_SST_HL: case F_SST_LL: {
      if ((rc = rec("sst", 1, a))) return 1;
      const double deltaT = kind == F_SST_HL ? -16.4 : 2.9; /* ocean_component.cpp:291, 300 */
      for (size_t q = 0; q < N; ++q) out[q] = a[q] + 18.0 + deltaT; /* MEAN_TOS_TEMP, oceanbox.hpp:35 */
      break;
    }
    case F_DIC_HL: case F_DIC_LL: {
      if ((rc = rec(kind == F_DIC_HL ? "HL_ocean_c" : "LL_ocean_c", 0, a))) return 1;
      const double vol = kind == F_DIC_HL ? C.vol_HL : C.vol_LL;
      for (size_t q = 0; q < N; ++q) out[q] = to_dic(a[q], vol);
      break;
    }
    case F_DIC:
      if ((rc = rec("LL_ocean_c", 0, a)) || (rc = rec("HL_ocean_c", 0, b))) return 1;
      for (size_t q = 0; q < N; ++q)
        out[q] = part_low * to_dic(a[q], C.vol_LL) + part_high * to_dic(b[q], C.vol_HL);
      break;
    case F_PH: case F_PCO2:
      if ((rc = rec(kind == F_PH ? "LL_pH" : "LL_PCO2", 0, a)) ||
          (rc = rec(kind == F_PH ? "HL_pH" : "HL_PCO2", 0, b)))
        return 1;
      for (size_t q = 0; q < N; ++q) out[q] = part_low * a[q] + part_high * b[q];
      break;
    case F_ML:
      if ((rc = rec("LL_ocean_c", 0, a)) || (rc = rec("HL_ocean_c", 0, b))) return 1;
      for (size_t q = 0; q < N; ++q) out[q] = a[q] + b[q];
      break;
    case F_OMEGACA_HL: case F_OMEGACA_LL: case F_OMEGAAR_HL: case F_OMEGAAR_LL:
    case F_CO3_HL: case F_CO3_LL: case F_CO3: {
      /* carbonate of the year's last chemistry solve, from what that solve left on record:
       * PCO2o = [CO2*] 1e6 / Kh and pH = -log10 h give [CO2*] and h; with K1, K2 at the box
       * temperature DIC = [CO2*] (1 + K1/h + K1 K2/h^2) and CO3 = DIC / (1 + h/K2 + h^2/(K1 K2))
       * (ocean_csys.cpp:166-366; Kh Weiss 1974, K1 / K2 Mehrbach refit) */
      std::vector<double> sst, pco2[2], ph[2];
      const double S = C.S;
      if ((rc = rec("sst", 1, sst))) return 1;
      const bool omega = kind >= F_OMEGACA_HL;
      const bool need_hl = kind == F_CO3_HL || kind == F_CO3 || kind == F_OMEGACA_HL || kind == F_OMEGAAR_HL;
      const bool need_ll = kind == F_CO3_LL || kind == F_CO3 || kind == F_OMEGACA_LL || kind == F_OMEGAAR_LL;
      /* saturation states: [CO3] [Ca] / Ksp with calcite's or aragonite's solubility product
       * (Mucci 1983; ocean_csys.cpp:300-318, 362-364) */
      auto omega_of = [&](double Tc, double co3_umol) {
        const double Tk = Tc + 273.15;
        const bool ca = kind == F_OMEGACA_HL || kind == F_OMEGACA_LL;
        double t1, t2, t3;
        if (ca) {
          t1 = -171.9065 - 0.077993 * Tk + 2839.319 / Tk + 71.595 * std::log10(Tk);
          t2 = +(-0.77712 + 0.0028426 * Tk + 178.34 / Tk) * std::sqrt(S);
          t3 = -0.07711 * S + 0.0041249 * std::pow(S, 1.5);
        } else {
          t1 = -171.945 - 0.077993 * Tk + 2903.293 / Tk + 71.595 * std::log10(Tk);
          t2 = +(-0.068393 + 0.0017276 * Tk + 88.135 / Tk) * std::sqrt(S);
          t3 = -0.10018 * S + 0.0059415 * std::pow(S, 1.5);
        }
        const double Ksp = std::pow(10.0, t1 + t2 + t3);
        const double calcium = 0.02128 / 40.087 * (S / 1.80655);
        return ((co3_umol / 1e6 * calcium) / Ksp);
      };
      if (need_hl && ((rc = rec("HL_PCO2", 0, pco2[0])) || (rc = rec("HL_pH", 0, ph[0])))) return 1;
      if (need_ll && ((rc = rec("LL_PCO2", 0, pco2[1])) || (rc = rec("LL_pH", 0, ph[1])))) return 1;
      auto co3 = [&](double Tc, double PCO2o, double pH) {
        const double Tk = Tc + 273.15;
        const double tmp = 9345.17 / Tk - 60.2409 + 23.3585 * std::log(Tk / 100);
        const double Kh = std::exp(tmp + S * (0.023517 - 0.00023656 * Tk + 0.0047036e-4 * Tk * Tk));
        const double K1 = std::pow(10, -(3633.86 / Tk - 61.2172 + 9.6777 * std::log(Tk) - 0.011555 * S + 0.0001152 * S * S));
        const double K2 = std::pow(10.0, -(471.78 / Tk + 25.9290 - 3.16967 * std::log(Tk) - 0.01781 * S + 0.0001122 * S * S));
        const double h = std::pow(10.0, -pH);
        const double co2st = PCO2o * Kh / 1e6;
        const double dic = co2st * (1.0 + K1 / h + K1 * K2 / h / h);
        return dic / (1.0 + h / K2 + h * h / K1 / K2) * 1e6;
      };
      for (size_t q = 0; q < N; ++q) {
        const double hl = need_hl ? co3(sst[q] + 18.0 + -16.4, pco2[0][q], ph[0][q]) : 0.0;
        const double ll = need_ll ? co3(sst[q] + 18.0 + 2.9, pco2[1][q], ph[1][q]) : 0.0;
        if (omega) out[q] = need_hl ? omega_of(sst[q] + 18.0 + -16.4, hl) : omega_of(sst[q] + 18.0 + 2.9, ll);
        else out[q] = kind == F_CO3_HL ? hl : kind == F_CO3_LL ? ll : part_low * ll + part_high * hl;
      }
      break;
    }
    case F_TAU_OH: {
      std::vector<double> M0, TOH0, CCH4, CNOX, CCO, CNMVOC;
      if ((rc = rec("CH4_concentration", 1, a)) || (rc = param(PI_M0, M0)) || (rc = param(PI_TOH0, TOH0)) ||
          (rc = param(PI_CCH4, CCH4)) || (rc = param(PI_CNOX, CNOX)) || (rc = param(PI_CCO, CCO)) ||
          (rc = param(PI_CNMVOC, CNMVOC)))
        return 1;
      for (int i = 0; i < M; ++i) {
        const double *R = raw[member_scen[i]].data();
        for (int k = 0; k < n_dates; ++k) {
          const int r = (int)dates[k] - cfg.start_year;
          const double previous_ch4 = a[(size_t)i * n_dates + k];
          double toh = 0.0;
          if (previous_ch4 != M0[i]) {
            const double ta = CCH4[i] * ((1.0 * std::log(previous_ch4)) - std::log(M0[i]));
            const double tb = CNOX[i] * ((1.0 * R[(size_t)RAW_NOX * nrow + r]) - R[(size_t)RAW_NOX * nrow]);
            const double tc = CCO[i] * ((1.0 * R[(size_t)RAW_CO * nrow + r]) - R[(size_t)RAW_CO * nrow]);
            const double td = CNMVOC[i] * ((1.0 * R[(size_t)RAW_NMVOC * nrow + r]) - R[(size_t)RAW_NMVOC * nrow]);
            toh = ta + tb + tc + td;
          }
          out[(size_t)i * n_dates + k] = TOH0[i] * std::exp(-toh);
        }
      }
      break;
    }
    case F_FROZEN: {
      /* Per biome the fraction only moves while that biome has permafrost left, so every year up
       * to the last one asked for is walked.  The datum without a biome prefix is the reference's
       * weighted mean (simpleNbox.cpp:492-513): the biomes' fractions of the DATE weighted with
       * their permafrost of the CURRENT date, and 1 when there is none in the system now. */
      const int nbm = n_biomes > 1 ? n_biomes : 1;
      std::vector<double> yrs(rmax + 1), T, now((size_t)M, 0.0);
      for (int r = 0; r <= rmax; ++r) yrs[r] = cfg.start_year + r;
      T.resize((size_t)M * rmax);
      if ((rc = hx_fetch(self, "land_tas", yrs.data(), rmax, T.data()))) return 1;
      const double root_two = 1.41421356237309504880168872420969807856967187537694;
      const double today = cfg.start_year + cur_row;
      std::vector<std::vector<double>> wnow(nbm, std::vector<double>((size_t)M)), fb(nbm, std::vector<double>(N));
      std::vector<double> P((size_t)M * rmax), wf((size_t)M), mu((size_t)M), sigma((size_t)M), f(rmax + 1);
      for (int ib = 0; ib < nbm; ++ib) {
        const std::string pre = n_biomes > 1 ? biome_names[ib] + "." : std::string();
        if ((rc = hx_fetch(self, (pre + "permafrost_c").c_str(), yrs.data(), rmax, P.data())) ||
            (rc = hx_fetch(self, (pre + "permafrost_c").c_str(), &today, 1, wnow[ib].data())) ||
            (rc = hx_get_param(self, (pre + "warmingfactor").c_str(), wf.data(), M)) ||
            (rc = hx_get_param(self, (pre + "pf_mu").c_str(), mu.data(), M)) ||
            (rc = hx_get_param(self, (pre + "pf_sigma").c_str(), sigma.data(), M)))
          return 1;
        for (int i = 0; i < M; ++i) {
          f[0] = 1.0;
          for (int r = 1; r <= rmax; ++r) { /* slowparameval of year r sees the year before */
            const double Tb = T[(size_t)i * rmax + r - 1] * wf[i], perm = P[(size_t)i * rmax + r - 1];
            f[r] = f[r - 1];
            if (perm == perm && perm != 0.0) {
              double cur = 1.0;
              if (Tb > 0) cur = 1 - std::erfc(-((std::log(Tb) - mu[i]) / (sigma[i] * root_two))) / 2;
              f[r] = cur;
            } else if (!(perm == perm)) f[r] = perm;
          }
          for (int k = 0; k < n_dates; ++k) fb[ib][(size_t)i * n_dates + k] = f[(int)dates[k] - cfg.start_year];
          now[i] += wnow[ib][i];
        }
      }
      for (int i = 0; i < M; ++i)
        for (int k = 0; k < n_dates; ++k) {
          const size_t at = (size_t)i * n_dates + k;
          double v = 1.0; /* no permafrost in the system now */
          if (one_biome >= 0) v = fb[one_biome][at];
          else if (now[i] > 0.0) {
            v = 0.0;
            for (int ib = 0; ib < nbm; ++ib) v += (wnow[ib][i] / now[i]) * fb[ib][at];
          } else if (!(now[i] == now[i])) v = now[i];
          out[at] = v;
        }
      break;
    }
    default: break;
  }
  /* a member that stopped reports nothing from its failing year on */
  std::vector<int32_t> st(M), fy(M);
  if ((rc = hx_member_status(self, st.data(), fy.data(), M))) return 1;
  for (int i = 0; i < M; ++i)
    if (st[i] > 0)
      for (int k = 0; k < n_dates; ++k)
        if (fy[i] <= (int)dates[k]) out[(size_t)i * n_dates + k] = std::nan("");
  return 1;
}

using hx::kParams;

#ifndef HX_REORDER_MAX_OUTPUTS
#define HX_REORDER_MAX_OUTPUTS 20
#endif

static int ensure_copy_stream(hx_engine *h) {
  if (h->copy_stream) return HX_OK;
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_seg, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess)
    return h->fail(HX_ERR_CUDA, "could not create the copy stream");
  return HX_OK;
}

extern "C" {

const char *hx_version(void) { return "hector_b200 0.1 (sm_100a)"; }

const char *hx_last_error(hx_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void hx_set_create_error(const char *msg) { g_create_error = msg ? msg : ""; }

int hx_create(const hx_config *cfg, hx_handle *out) {
  if (!cfg || !out) {
    g_create_error = "hx_create: null argument";
    return HX_ERR_ARG;
  }
  *out = nullptr;
  if (cfg->n_members <= 0 || cfg->n_scenarios <= 0 || cfg->end_year <= cfg->start_year) {
    g_create_error = "hx_create: need n_members > 0, n_scenarios > 0, end_year > start_year";
    return HX_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("hx_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); the engine has no CPU path";
    return HX_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "hx_create: bad device ordinal";
    return HX_ERR_ARG;
  }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return HX_ERR_CUDA;
  }
  hx_engine *h = new hx_engine();
  h->cfg = *cfg;
  h->M = cfg->n_members;
  h->nscen = cfg->n_scenarios;
  h->nrow = cfg->end_year - cfg->start_year + 1;
  h->raw.assign(h->nscen, std::vector<double>((size_t)RAW_COUNT * h->nrow, 0.0));
  h->raw_set.assign(h->nscen, std::vector<char>(RAW_COUNT, 0));
  h->cons.assign(h->nscen, std::vector<double>((size_t)CN_COUNT * h->nrow, NAN));
  h->member_scen.assign(h->M, 0);
  for (int i = 0; i < PI_COUNT; ++i) {
    h->pscalar[i] = kParams[i].dflt;
    h->pvec_on_device_only[i] = false;
  }
  for (int g = 0; g < HX_NHALO; ++g) {
    h->halo_tau[g] = hx::kHaloTau[g]; h->halo_rho[g] = hx::kHaloRho[g];
    h->halo_delta[g] = hx::kHaloDelta[g]; h->halo_H0[g] = hx::kHaloH0[g];
    h->halo_mm[g] = hx::kHaloMolarMass[g];
  }
  h->out_sel = {OUT_CO2, OUT_TAS};
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    g_create_error = "hx_create: could not create stream/events";
    delete h;
    return HX_ERR_CUDA;
  }
  h->stream = h->own_stream;
  *out = h;
  return HX_OK;
}

int hx_destroy(hx_handle h) {
  if (!h) return HX_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->free_device();
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->h_slab_done) cudaFreeHost(h->h_slab_done);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  hx_ipc_close(h);
  hx_xchg_close(h);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_rec_full[i]) cudaEventDestroy(h->ev_rec_full[i]);
    if (h->ev_rec_free[i]) cudaEventDestroy(h->ev_rec_free[i]);
  }
  if (h->track_stream) cudaStreamDestroy(h->track_stream);
  for (cudaEvent_t e : h->ev_user)
    if (e) cudaEventDestroy(e);
  if (h->ev_seg) cudaEventDestroy(h->ev_seg);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->ev_setup_a) cudaEventDestroy(h->ev_setup_a);
  if (h->ev_setup_b) cudaEventDestroy(h->ev_setup_b);
  if (h->setup_stream) cudaStreamDestroy(h->setup_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return HX_OK;
}

int hx_set_stream(hx_handle h, void *cuda_stream) {
  if (!h) return HX_ERR_ARG;
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return HX_OK;
}

int hx_set_scenario_series(hx_handle h, int32_t scenario_id, const char *name, int32_t year0,
                           int32_t n, const double *values) {
  if (!h || !name || !values) return HX_ERR_ARG;
  if (scenario_id < 0 || scenario_id >= h->nscen) return h->fail(HX_ERR_ARG, "bad scenario id");
  const int ci = Engine::find_constraint(name);
  if (ci >= 0) {
    if (h->prepared && h->d_GP && (ci == CN_N2O || ci >= CN_HALO0))
      return h->fail(HX_ERR_UNSUPPORTED, "an N2O or halocarbon concentration constraint cannot be added to a "
                                         "run with per-member N2O / halocarbon parameters");
    /* a constraint series may cover any part of the run; NaN = no entry for that year */
    double *dst = h->cons[scenario_id].data() + (size_t)ci * h->nrow;
    for (int k = 0; k < n; ++k) { /* entries outside [year0, year0 + n) are kept */
      const int r = year0 + k - h->cfg.start_year;
      if (r >= 0 && r < h->nrow) dst[r] = values[k];
    }
    if (h->prepared) {
      h->tables_dirty = true; h->params_dirty = true;
      h->dirty_from_row = std::min(h->dirty_from_row, std::max(0, year0 - h->cfg.start_year));
    }
    return HX_OK;
  }
  const int si = Engine::find_raw(name);
  if (si < 0) return h->fail(HX_ERR_ARG, std::string("unknown scenario series: ") + name);
  if (h->prepared && h->d_GP && (si == RAW_N2O_E || si == RAW_N2O_NAT || si >= RAW_HALO0))
    return h->fail(HX_ERR_UNSUPPORTED, "with per-member N2O / halocarbon parameters the gas emission "
                                       "series are fixed at hx_prepare");

  double *dst = h->raw[scenario_id].data() + (size_t)si * h->nrow;
  if (year0 > h->cfg.start_year || year0 + n - 1 < h->cfg.end_year) {
    /* a few years of a series that is already there: R's setvar(core, dates, var, values)
     * (R/messages.R:107-140 -> Core::sendMessage(SETDATA, var, message_data(date, value)));
     * a series must have been given in full once (the ini / csv reader does that) */
    if (!h->raw_set[scenario_id][si])
      return h->fail(HX_ERR_ARG, std::string("series does not cover start..end: ") + name);
    for (int k = 0; k < n; ++k) {
      const int r = year0 + k - h->cfg.start_year;
      if (r >= 0 && r < h->nrow && values[k] == values[k]) dst[r] = values[k];
    }
  } else {
    for (int r = 0; r < h->nrow; ++r) dst[r] = values[h->cfg.start_year - year0 + r];
    h->raw_set[scenario_id][si] = 1;
  }
  if (h->prepared) { /* R: setvar + reset */
    h->tables_dirty = true; h->params_dirty = true;
    h->dirty_from_row = std::min(h->dirty_from_row, std::max(0, year0 - h->cfg.start_year));
  }
  return HX_OK;
}

int hx_set_scenario_table(hx_handle h, int32_t scenario_id, int32_t n_names,
                          const char *const *names, int32_t year0, int32_t n_years,
                          const double *values) {
  if (!h || !names || !values) return HX_ERR_ARG;
  std::vector<double> col(n_years);
  for (int j = 0; j < n_names; ++j) {
    for (int r = 0; r < n_years; ++r) col[r] = values[(size_t)r * n_names + j];
    int rc = hx_set_scenario_series(h, scenario_id, names[j], year0, n_years, col.data());
    if (rc) return rc;
  }
  return HX_OK;
}

int hx_set_member_scenario(hx_handle h, const int32_t *scen, int32_t n) {
  if (!h || !scen) return HX_ERR_ARG;
  if (h->prepared) return h->fail(HX_ERR_STATE, "member scenarios must be set before hx_prepare");
  if (n != h->M) return h->fail(HX_ERR_ARG, "hx_set_member_scenario: n != n_members");
  for (int i = 0; i < n; ++i)
    if (scen[i] < 0 || scen[i] >= h->nscen) return h->fail(HX_ERR_ARG, "scenario id out of range");
  h->member_scen.assign(scen, scen + n);
  return HX_OK;
}

static int set_special_scalar(hx_engine *h, const char *name, double v) {
  {
    const int gp = Engine::find_gas_param(name);
    if (gp >= 0) h->gvec[gp].clear();
  }
  if (!strcmp(name, "trackingDate")) { h->tracking_date = (int)v; return 1; }
  if (!strcmp(name, "baseyear")) { h->baseyear = v; return 1; }
  if (!strcmp(name, "max_spinup")) { h->max_spinup = (int)v; return 1; }
  if (!strcmp(name, "UC_N2O")) { h->UC_N2O = v; return 1; }
  if (!strcmp(name, "TN2O0")) { h->TN2O0 = v; return 1; }
  const char *dot = strchr(name, '.');
  if (dot) { /* "<gas>.tau" etc. */
    std::string gas(name, dot - name), field(dot + 1);
    for (int g = 0; g < HX_NHALO; ++g)
      if (gas == hx::kHaloNames[g]) {
        if (field == "tau") h->halo_tau[g] = v;
        else if (field == "rho") h->halo_rho[g] = v;
        else if (field == "delta") h->halo_delta[g] = v;
        else if (field == "H0") h->halo_H0[g] = v;
        else if (field == "molarMass") h->halo_mm[g] = v;
        else return 0;
        return 1;
      }
  }
  return 0;
}

int hx_set_biomes(hx_handle h, int32_t n_biomes, const char *const *names) {
  if (!h) return HX_ERR_ARG;
  if (h->prepared) return h->fail(HX_ERR_STATE, "hx_set_biomes must be called before hx_prepare");
  if (n_biomes <= 1 && !names) { /* back to the single global biome */
    h->n_biomes = 1;
    h->biome_names.clear();
    return HX_OK;
  }
  if (n_biomes < 2 || n_biomes > HX_MAX_BIOMES || !names)
    return h->fail(HX_ERR_ARG, "hx_set_biomes: 2 .. " + std::to_string(HX_MAX_BIOMES) + " named biomes");
  std::vector<std::string> nm;
  for (int i = 0; i < n_biomes; ++i) {
    if (!names[i] || !*names[i] || strchr(names[i], '.') || !strcmp(names[i], "global"))
      return h->fail(HX_ERR_ARG, "hx_set_biomes: bad biome name (empty, dotted or 'global')");
    for (const std::string &o : nm)
      if (o == names[i]) return h->fail(HX_ERR_ARG, std::string("biome listed twice: ") + names[i]);
    nm.push_back(names[i]);
  }
  h->n_biomes = n_biomes;
  h->biome_names = nm;
  /* per-biome outputs selected under an earlier biome list no longer mean anything */
  h->out_sel.erase(std::remove_if(h->out_sel.begin(), h->out_sel.end(),
                                  [](int id) { return id >= OUT_COUNT; }),
                   h->out_sel.end());
  for (int ib = 0; ib < HX_MAX_BIOMES; ++ib)
    for (int f = 0; f < BP_COUNT; ++f) {
      h->bscalar[ib][f] = hx::kBiomeParams[f].dflt;
      h->bvec[ib][f].clear();
    }
  return HX_OK;
}

int hx_biome_count(hx_handle h) { return h ? h->n_biomes : HX_ERR_ARG; }

int hx_biome_name(hx_handle h, int32_t i, char *buf, int32_t cap) {
  if (!h || !buf || cap < 1) return HX_ERR_ARG;
  const std::string nm = h->n_biomes <= 1 ? (i == 0 ? "global" : "") : (i >= 0 && i < h->n_biomes ? h->biome_names[i] : "");
  if (nm.empty()) return h->fail(HX_ERR_ARG, "hx_biome_name: index out of range");
  snprintf(buf, (size_t)cap, "%s", nm.c_str());
  return HX_OK;
}

int hx_tracking_date(hx_handle h) { return h ? h->tracking_date : HX_ERR_ARG; }

/* "<biome>.<name>" inputs: scalar (per_member null) or one value per member */
static int set_biome_param(hx_handle h, int ib, int f, double value, const double *per_member) {
  if (per_member) h->bvec[ib][f].assign(per_member, per_member + h->M);
  else { h->bscalar[ib][f] = value; h->bvec[ib][f].clear(); }
  if (h->prepared) { /* like a global parameter: takes effect at the next reset / run */
    cudaSetDevice(h->cfg.device);
    int rc = h->upload_biomes();
    if (rc) return rc;
    h->params_dirty = true;
    h->dirty_from_row = 0;
    h->spinup_dirty = true;
  }
  return HX_OK;
}

int hx_set_param_scalar(hx_handle h, const char *name, double value) {
  if (!h || !name) return HX_ERR_ARG;
  {
    int ib, f;
    if (h->find_biome_param(name, ib, f)) return set_biome_param(h, ib, f, value, nullptr);
  }
  const int pi = h->find_param(name);
  if (pi >= 0 && h->n_biomes > 1 && Engine::is_biome_replaced(pi))
    return h->fail(HX_ERR_ARG, std::string(name) + ": cannot have both global and biome-specific "
                                                   "data (simpleNbox-runtime.cpp:66-69)");
  if (pi < 0) {
    if (h->prepared) return h->fail(HX_ERR_STATE, std::string(name) + " must be set before hx_prepare");
    if (set_special_scalar(h, name, value)) return HX_OK;
    return h->fail(HX_ERR_ARG, std::string("unknown parameter: ") + name);
  }
  if (pi == PI_N0 && h->prepared)
    return h->fail(HX_ERR_STATE, "N0 feeds the host N2O series; set it before hx_prepare");
  h->pscalar[pi] = value;
  h->pvec[pi].clear();
  h->pvec_on_device_only[pi] = false;
  if (h->prepared) {
    cudaSetDevice(h->cfg.device);
    int rc = h->upload_param(pi);
    if (rc) return rc;
    h->params_dirty = true;
    h->dirty_from_row = 0;
    if (Engine::affects_spinup(pi)) h->spinup_dirty = true;
  }
  return HX_OK;
}

int hx_set_param(hx_handle h, const char *name, const double *per_member, int32_t n) {
  if (!h || !name || !per_member) return HX_ERR_ARG;
  {
    int ib, f;
    if (h->find_biome_param(name, ib, f)) {
      if (n != h->M) return h->fail(HX_ERR_ARG, "hx_set_param: n != n_members");
      return set_biome_param(h, ib, f, 0.0, per_member);
    }
  }
  const int pi = h->find_param(name);
  if (pi >= 0 && h->n_biomes > 1 && Engine::is_biome_replaced(pi))
    return h->fail(HX_ERR_ARG, std::string(name) + ": cannot have both global and biome-specific "
                                                   "data (simpleNbox-runtime.cpp:66-69)");
  if (pi < 0) {
    /* N2O / halocarbon parameters per member: the run then carries the 27 gas recurrences on
     * the device (n2o_component.cpp:98-116, halocarbon_component.cpp:127-136) */
    const int gp = Engine::find_gas_param(name);
    if (gp < 0) return h->fail(HX_ERR_ARG, std::string("unknown per-member parameter: ") + name);
    if (h->prepared) return h->fail(HX_ERR_STATE, std::string(name) + " per member must be set before hx_prepare");
    if (n != h->M) return h->fail(HX_ERR_ARG, "hx_set_param: n != n_members");
    h->gvec[gp].assign(per_member, per_member + n);
    return HX_OK;
  }
  if (pi == PI_N0 && h->prepared)
    return h->fail(HX_ERR_STATE, "N0 per member must be set before hx_prepare");
  if (n != h->M) return h->fail(HX_ERR_ARG, "hx_set_param: n != n_members");
  h->pvec[pi].assign(per_member, per_member + n);
  h->pvec_on_device_only[pi] = false;
  if (h->prepared) {
    cudaSetDevice(h->cfg.device);
    int rc = h->upload_param(pi);
    if (rc) return rc;
    h->params_dirty = true;
    h->dirty_from_row = 0;
    if (Engine::affects_spinup(pi)) h->spinup_dirty = true;
  }
  return HX_OK;
}

/* one member's value of a per-member parameter (Core::sendMessage(M_SETDATA) addressed to one
 * core of an ensemble): the host copy and, once prepared, the one device element */
int hx_set_param_member(hx_handle h, const char *name, int32_t member, double value) {
  if (!h || !name) return HX_ERR_ARG;
  if (member < 0 || member >= h->M) return h->fail(HX_ERR_ARG, "member index out of range");
  int ib, f;
  const int pi = h->find_param(name);
  const bool general = h->find_biome_param(name, ib, f) || pi < 0 || pi == PI_N0 || h->pvec_on_device_only[pi] ||
                       (h->n_biomes > 1 && Engine::is_biome_replaced(pi));
  if (general) { /* biome / gas parameters, device-resident vectors: through the whole vector */
    std::vector<double> cur((size_t)h->M);
    int rc = hx_get_param(h, name, cur.data(), h->M);
    if (rc) return rc;
    cur[member] = value;
    return hx_set_param(h, name, cur.data(), h->M);
  }
  if (h->pvec[pi].empty()) h->pvec[pi].assign((size_t)h->M, h->pscalar[pi]);
  h->pvec[pi][member] = value;
  if (h->prepared) {
    cudaSetDevice(h->cfg.device);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy(h->d_P + HX_TILED(pi, h->dev_of_api[member], PD_COUNT), &value, sizeof(double),
                     cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_set_param_member: ") + cudaGetErrorString(e));
    h->params_dirty = true;
    h->dirty_from_row = 0;
    if (Engine::affects_spinup(pi)) h->spinup_dirty = true;
  }
  return HX_OK;
}

int hx_set_param_device(hx_handle h, const char *name, const double *dev, int32_t n) {
  if (!h || !name || !dev) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_set_param_device needs hx_prepare first");
  const int pi = h->find_param(name);
  if (pi < 0 || pi == PI_N0) return h->fail(HX_ERR_ARG, std::string("bad per-member parameter: ") + name);
  if (n != h->M) return h->fail(HX_ERR_ARG, "hx_set_param_device: n != n_members");
  cudaSetDevice(h->cfg.device);
  k_scatter_field<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_P, pi, PD_COUNT, dev, h->d_dev_of_api, n);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  }
  h->pvec[pi].clear();
  h->pvec_on_device_only[pi] = true;
  h->params_dirty = true;
  h->dirty_from_row = 0;
  if (Engine::affects_spinup(pi)) h->spinup_dirty = true;
  return HX_OK;
}

int hx_get_param(hx_handle h, const char *name, double *out, int32_t n) {
  if (!h || !name || !out) return HX_ERR_ARG;
  {
    int ib, f;
    if (h->find_biome_param(name, ib, f)) { /* "<biome>.<name>": host copies (set before prepare) */
      if (n != h->M) return h->fail(HX_ERR_ARG, "hx_get_param: n != n_members");
      for (int i = 0; i < n; ++i)
        out[i] = h->bvec[ib][f].empty() ? h->bscalar[ib][f] : h->bvec[ib][f][i];
      return HX_OK;
    }
  }
  if (!strcmp(name, "baseyear")) { /* forcing_component.cpp getData(D_RF_BASEYEAR) */
    if (n != h->M) return h->fail(HX_ERR_ARG, "hx_get_param: n != n_members");
    for (int i = 0; i < n; ++i) out[i] = h->baseyear == 0.0 ? h->cfg.start_year + 1 : h->baseyear;
    return HX_OK;
  }
  const int pi = h->find_param(name);
  if (pi < 0) {
    const int gp = Engine::find_gas_param(name);
    if (gp < 0) return h->fail(HX_ERR_ARG, std::string("unknown parameter: ") + name);
    if (n != h->M) return h->fail(HX_ERR_ARG, "hx_get_param: n != n_members");
    for (int i = 0; i < n; ++i) out[i] = h->gas_value(gp, i);
    return HX_OK;
  }
  if (n != h->M) return h->fail(HX_ERR_ARG, "hx_get_param: n != n_members");
  if (h->pvec_on_device_only[pi]) {
    cudaSetDevice(h->cfg.device);
    if (h->ensure_stage((size_t)n * sizeof(double))) return HX_ERR_CUDA;
    k_gather_field<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_stage, h->d_P, pi, PD_COUNT,
                                                          h->d_dev_of_api, n);
    cudaError_t e = cudaMemcpyAsync(out, h->d_stage, (size_t)n * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  } else if (h->pvec[pi].empty()) {
    for (int i = 0; i < n; ++i) out[i] = h->pscalar[pi];
  } else {
    memcpy(out, h->pvec[pi].data(), (size_t)n * sizeof(double));
  }
  return HX_OK;
}

int hx_select_outputs(hx_handle h, int32_t n, const char *const *names) {
  if (!h || (n > 0 && !names)) return HX_ERR_ARG;
  if (h->prepared) return h->fail(HX_ERR_STATE, "outputs must be selected before hx_prepare");
  std::vector<int> sel;
  for (int i = 0; i < n; ++i) {
    const int id = h->find_out(names[i]);
    if (id < 0) return h->fail(HX_ERR_ARG, std::string("unknown output variable: ") + names[i]);
    if (std::find(sel.begin(), sel.end(), id) == sel.end()) sel.push_back(id);
  }
  h->out_sel = sel;
  return HX_OK;
}

int hx_prepare(hx_handle h) {
  NvtxRange nvtx_("hx_prepare");
  if (!h) return HX_ERR_ARG;
  if (h->prepared) return h->fail(HX_ERR_STATE, "hx_prepare called twice");
  cudaSetDevice(h->cfg.device);
  for (int s = 0; s < h->nscen; ++s)
    for (int i = 0; i < RAW_COUNT; ++i)
      if (!h->raw_set[s][i]) {
        std::string nm = i < RAW_HALO0 ? hx::kRawNames[i]
                                       : std::string(hx::kHaloNames[i - RAW_HALO0]) + "_emissions";
        return h->fail(HX_ERR_STATE, "scenario " + std::to_string(s) + ": series not set: " + nm);
      }
  auto fail = [&](int code, const char *msg) { return h->fail(code, msg); };
  const int M = h->M, nrow = h->nrow;

  /* members grouped by scenario; every scenario segment padded to whole CTAs */
  std::vector<int> count(h->nscen, 0), seg_start(h->nscen, 0);
  for (int i = 0; i < M; ++i) count[h->member_scen[i]]++;
  int Mpad = 0;
  std::vector<int32_t> block_scen;
  for (int s = 0; s < h->nscen; ++s) {
    seg_start[s] = Mpad;
    const int nb = (count[s] + HX_BLOCK - 1) / HX_BLOCK;
    for (int b = 0; b < nb; ++b) block_scen.push_back(s);
    Mpad += nb * HX_BLOCK;
  }
  h->Mpad = Mpad;
  /* Inside a scenario the members are laid out so that the 32 members of a warp -- and the 128
   * of a tile -- behave alike.  Members differ in how many ODE sub-steps a year takes (1 to 4,
   * following the ocean's reduced-time-step machine) and a warp pays for its slowest lane: in
   * caller order an LHS ensemble keeps 26.6 of 32 lanes busy (ncu, round 1).  How a member
   * behaves is a smooth function of its parameters, so the members are sorted into a k-d tree
   * over the per-member parameters that vary (median splits, cycling through the parameters,
   * down to leaves of one warp): neighbours in every parameter share a warp.  Outputs stay in
   * API order (api_of_dev), so nothing outside the engine sees the permutation. */
  h->dev_of_api.resize(M);
  std::vector<int32_t> api_of_dev(Mpad, -1);
  {
    std::vector<std::vector<int>> of_scen(h->nscen);
    for (int i = 0; i < M; ++i) of_scen[h->member_scen[i]].push_back(i);
    std::vector<int> dims;
    static const int first_dims[] = {PI_BETA, PI_S, PI_DIFF, PI_Q10};
    for (int pi : first_dims)
      if (!h->pvec[pi].empty()) dims.push_back(pi);
    for (int pi = 0; pi < PI_COUNT; ++pi)
      if (!h->pvec[pi].empty() && std::find(dims.begin(), dims.end(), pi) == dims.end()) dims.push_back(pi);
    for (int s = 0; s < h->nscen; ++s) {
      std::vector<int> &ix = of_scen[s];
      /* ... at a price: the outputs are stored in API order, so a warp of re-ordered members
       * scatters every output row it writes over 32 sectors.  Two outputs cost 1.8 GB of extra
       * DRAM traffic against a 13 % faster year loop; with every output recorded the scattered
       * stores take over (65 536 members: 52.3 ms re-ordered against 33.6 ms in caller order,
       * the crossover near 20 outputs, tools/profile_outputs.py), so runs that record more than
       * HX_REORDER_MAX_OUTPUTS variables keep the caller's order. */
      const bool reorder = !(h->cfg.flags & HX_FLAG_KEEP_ORDER) && (int)h->out_sel.size() <= HX_REORDER_MAX_OUTPUTS;
      if (!dims.empty() && reorder) {
        /* iterative k-d ordering: (lo, hi, depth) ranges split at the median of one parameter */
        struct Range { int lo, hi, depth; };
        std::vector<Range> todo(1, Range{0, (int)ix.size(), 0});
        while (!todo.empty()) {
          const Range r = todo.back();
          todo.pop_back();
          if (r.hi - r.lo <= 32) continue;
          const std::vector<double> &v = h->pvec[dims[r.depth % dims.size()]];
          /* split at a multiple of 32 so that leaves are whole warps */
          int mid = r.lo + ((r.hi - r.lo) / 2 + 31) / 32 * 32;
          if (mid >= r.hi) mid = r.lo + (r.hi - r.lo) / 2;
          std::nth_element(ix.begin() + r.lo, ix.begin() + mid, ix.begin() + r.hi,
                           [&](int a, int b) { return v[a] < v[b] || (v[a] == v[b] && a < b); });
          todo.push_back(Range{r.lo, mid, r.depth + 1});
          todo.push_back(Range{mid, r.hi, r.depth + 1});
        }
      }
      for (size_t k = 0; k < ix.size(); ++k) {
        h->dev_of_api[ix[k]] = seg_start[s] + (int)k;
        api_of_dev[seg_start[s] + (int)k] = ix[k];
      }
    }
  }
  std::vector<int32_t> status(Mpad, -1);
  for (int i = 0; i < M; ++i) status[h->dev_of_api[i]] = 0;
  h->first_active = h->dev_of_api[0];

  /* constants */
  HxConst &C = h->C;
  C.start_year = h->cfg.start_year; C.end_year = h->cfg.end_year; C.nrow = nrow;
  C.baseyear = h->baseyear == 0.0 ? h->cfg.start_year + 1 : (int)h->baseyear;
  C.max_spinup = h->max_spinup;
  C.flags = h->cfg.flags;
  C.S = 34.5; /* ocean_component.cpp:291 */
  C.sqrtS = std::sqrt(C.S);
  C.S15 = std::pow(C.S, (3.0 / 2.0));
  C.bor = 1 * (416.0 * (C.S / 35.0)) * 1.e-6; /* ocean_csys.cpp:287 */
  {
    const double part_high = 0.15, part_low = 1 - part_high;
    const double thick_LL = 100, thick_HL = 100, thick_inter = 1000 - thick_LL,
                 thick_deep = 3777 - thick_inter - thick_LL;
    const double ocean_area = 3.6e14;
    C.vol_LL = ocean_area * part_low * thick_LL;
    C.vol_HL = ocean_area * part_high * thick_HL;
    C.inv_vol_LL = 1.0 / C.vol_LL;
    C.inv_vol_HL = 1.0 / C.vol_HL;
    C.vol_IO = ocean_area * thick_inter;
    C.vol_DO = ocean_area * thick_deep;
    C.As_HL = ocean_area * part_high;
    C.As_LL = ocean_area * part_low;
    C.U = 6.7;
    C.spy_ocean = 60 * 60 * 24 * 365.25;
  }
  {
    const double ocean_area = (1.0 - 0.29) * 5100656E8;
    C.powtoheat = ocean_area * (60.0 * 60.0 * 24.0 * 365.2422) / std::pow(10.0, 22);
  }
  C.rk_grow_max = 9.0 / 10.0 * std::pow(std::pow(5.0, -5.0), -1.0 / 5.0);

  /* carbon tracking: recorded years */
  const bool tracking = h->tracking_date <= h->cfg.end_year;
  h->track_years.clear();
  if (tracking) {
    if (h->tracking_date <= h->cfg.start_year)
      return fail(HX_ERR_ARG, "trackingDate must lie in (startDate, endDate]"); /* core.cpp:228-235 */
    if (h->track_every > 0)
      for (int y = h->tracking_date; y <= h->cfg.end_year; y += h->track_every) h->track_years.push_back(y);
    if (h->track_years.empty() || h->track_years.back() != h->cfg.end_year)
      h->track_years.push_back(h->cfg.end_year);
  }
  C.tracking_date = tracking ? h->tracking_date : 0x7fffffff;
  C.track_every = h->track_every;
  C.track_nrec = (int)h->track_years.size();

  /* biomes: every biome needs its pools and parameters (simpleNbox-runtime.cpp:66-101) */
  const int nb = h->n_biomes;
  C.n_biomes = nb;
  for (int i = 0; i < HX_MAX_BIOMES; ++i) C.biome_order[i] = i;
  if (nb > 1) {
    if (tracking)
      return fail(HX_ERR_UNSUPPORTED, "carbon tracking with more than one biome is not implemented");
    for (int ib = 0; ib < nb; ++ib)
      for (int f = 0; f < BP_COUNT; ++f)
        if (h->bvec[ib][f].empty() && std::isnan(h->bscalar[ib][f]))
          return fail(HX_ERR_ARG, ("no " + std::string(hx::kBiomeParams[f].name) + " data for biome " +
                                   h->biome_names[ib]).c_str());
    std::sort(C.biome_order, C.biome_order + nb,
              [&](int a, int b) { return h->biome_names[a] < h->biome_names[b]; });
  }

  /* device scenario tables */
  std::vector<double> tab;
  bool any_constraint = false, any_nbp = false;
  h->build_tables(tab, any_constraint, any_nbp);
  const bool gas = h->gas_per_member();
  std::vector<double> gas_tab;
  if (gas) {
    bool gas_constraint = false;
    for (int sc = 0; sc < h->nscen; ++sc)
      for (int series = CN_N2O; series < CN_COUNT; series = (series == CN_N2O ? CN_HALO0 : series + 1)) {
        const double *c = h->con(sc, series);
        for (int r = 0; r < nrow; ++r) gas_constraint = gas_constraint || c[r] == c[r];
      }
    if (gas_constraint)
      return fail(HX_ERR_UNSUPPORTED, "per-member N2O / halocarbon parameters are not combined with "
                                      "N2O / halocarbon concentration constraints");
    gas_tab.assign((size_t)h->nscen * nrow * HX_GAS_COLS, 0.0);
    for (int sc = 0; sc < h->nscen; ++sc) {
      const double *R = h->raw[sc].data();
      for (int r = 0; r < nrow; ++r) {
        double *row = gas_tab.data() + ((size_t)sc * nrow + r) * HX_GAS_COLS;
        row[0] = R[(size_t)RAW_N2O_E * nrow + r];
        row[1] = R[(size_t)RAW_N2O_NAT * nrow + r];
        for (int g = 0; g < HX_NHALO; ++g) row[2 + g] = R[(size_t)(RAW_HALO0 + g) * nrow + r];
      }
    }
  }

  const size_t Mp = Mpad;
  const int nsel = (int)h->out_sel.size();
  /* the per-stash outputs (hx_layout.h, "scratch rows X") */
  bool want_x = false;
  for (int id : h->out_sel) want_x = want_x || (id >= OUT_UPTAKE_HL && id <= OUT_RH_SOIL);
  if (tracking) {
    /* slabs per tracked launch (launch_rows): up to 4, from the memory the device has free --
     * two buffers of rec_group slab records, at most 45 % of it (HX_TRK_GROUP overrides) */
    size_t free_b = 0, total_b = 0;
    const size_t slab_b = block_scen.size() * hx::track_record_bytes_per_cta();
    int g = 1;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && slab_b > 0)
      g = (int)std::min<size_t>(4, std::max<size_t>(1, (size_t)(0.45 * (double)free_b) / (2 * slab_b)));
    if (const char *ev = std::getenv("HX_TRK_GROUP")) g = std::max(1, std::min(64, std::atoi(ev)));
    h->rec_group = g;
  }

  /* parameters and derived constants are one tiled array, [tile][PI_COUNT | DI_COUNT][128]: one
   * base pointer per thread serves both (tried: a persisting-L2 window over it against the slabs'
   * history streams -- 31.59 vs 31.65 ms, not kept) */
  if (cudaMalloc(&h->d_P, (size_t)PD_COUNT * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_S, SI_COUNT * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_S_snap, SI_COUNT * Mp * sizeof(double)) != cudaSuccess ||
      (h->d_D = h->d_P + (size_t)PI_COUNT * HX_TILE) == nullptr || /* tile 0's derived constants */
      cudaMalloc(&h->d_ker, (size_t)HX_KER_ROWS(nrow) * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_conv, (size_t)HX_SLAB_YEARS * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_sst, (size_t)nrow * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_tland, (size_t)nrow * Mp * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_out, std::max<size_t>(1, (size_t)nsel * (nrow - 1) * Mp) * sizeof(double)) != cudaSuccess ||
      (want_x && cudaMalloc(&h->d_X, (size_t)XS_COUNT * Mp * sizeof(double)) != cudaSuccess) ||
      cudaMalloc(&h->d_scen, tab.size() * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&h->d_block_scen, block_scen.size() * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_status, Mp * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_status_snap, Mp * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_status_post, Mp * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_fail_year, Mp * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_spinup_steps, Mp * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_counters, HX_NCOUNTERS * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&h->d_sched, (2 * block_scen.size() + 1 + nrow / HX_SLAB_YEARS + 2) * sizeof(unsigned)) != cudaSuccess ||
      cudaMalloc(&h->d_dev_of_api, (size_t)M * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&h->d_api_of_dev, Mp * sizeof(int32_t)) != cudaSuccess ||
      (gas &&
       (cudaMalloc(&h->d_GP, (size_t)GP_COUNT * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_GF, (size_t)GF_COUNT * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_GF_snap, (size_t)GF_COUNT * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_scen_gas, gas_tab.size() * sizeof(double)) != cudaSuccess)) ||
      (nb > 1 &&
       (cudaMalloc(&h->d_BP, (size_t)nb * BP_COUNT * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_BF, (size_t)nb * BF_COUNT * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_BF_snap, (size_t)nb * BF_COUNT * Mp * sizeof(double)) != cudaSuccess)) ||
      (tracking &&
       (cudaMalloc(&h->d_T, (size_t)TS_COUNT * HX_NSRC * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_TK, (size_t)TS_COUNT * Mp * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&h->d_TO, h->track_years.size() * HX_NPOOL * HX_NSRC * Mp * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&h->d_TOK, h->track_years.size() * HX_NPOOL * Mp * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&h->d_REC, 2 * (size_t)h->rec_group * block_scen.size() * hx::track_record_bytes_per_cta()) != cudaSuccess ||
        cudaMalloc(&h->d_YCNT, 2 * (size_t)h->rec_group * block_scen.size() * hx::track_ycnt_bytes_per_tile()) != cudaSuccess ||
        cudaMalloc(&h->d_trk_fail, Mp * sizeof(int32_t)) != cudaSuccess))) {
    cudaError_t e = cudaGetLastError();
    h->free_device();
    return fail(HX_ERR_CUDA, (std::string("device allocation failed: ") + cudaGetErrorString(e)).c_str());
  }
  cudaStream_t st = h->stream;
  {
    cudaError_t e = cudaMemcpyAsync(h->d_scen, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    auto also = [&e](cudaError_t r) { if (e == cudaSuccess) e = r; };
    also(cudaMemcpyAsync(h->d_block_scen, block_scen.data(), block_scen.size() * sizeof(int32_t),
                         cudaMemcpyHostToDevice, st));
    also(cudaMemcpyAsync(h->d_status_snap, status.data(), Mp * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    also(cudaMemcpyAsync(h->d_dev_of_api, h->dev_of_api.data(), (size_t)M * sizeof(int32_t),
                         cudaMemcpyHostToDevice, st));
    also(cudaMemcpyAsync(h->d_api_of_dev, api_of_dev.data(), Mp * sizeof(int32_t),
                         cudaMemcpyHostToDevice, st));
    also(cudaMemsetAsync(h->d_counters, 0, HX_NCOUNTERS * sizeof(unsigned long long), st));
    also(cudaMemsetAsync(h->d_fail_year, 0, Mp * sizeof(int32_t), st));
    also(cudaMemsetAsync(h->d_spinup_steps, 0, Mp * sizeof(int32_t), st));
    also(cudaMemsetAsync(h->d_S, 0, SI_COUNT * Mp * sizeof(double), st));
    also(cudaMemsetAsync(h->d_P, 0, (size_t)PD_COUNT * Mp * sizeof(double), st)); /* parameters follow below */
    also(cudaMemsetAsync(h->d_ker, 0, (size_t)HX_KER_ROWS(nrow) * Mp * sizeof(double), st));
    also(cudaMemsetAsync(h->d_conv, 0, (size_t)HX_SLAB_YEARS * Mp * sizeof(double), st));
    if (h->d_X) also(cudaMemsetAsync(h->d_X, 0, (size_t)XS_COUNT * Mp * sizeof(double), st));
    also(cudaMemsetAsync(h->d_sst, 0, (size_t)nrow * Mp * sizeof(double), st));
    also(cudaMemsetAsync(h->d_tland, 0, (size_t)nrow * Mp * sizeof(double), st));
    if (h->d_BF) also(cudaMemsetAsync(h->d_BF, 0, (size_t)nb * BF_COUNT * Mp * sizeof(double), st));
    if (gas) {
      also(cudaMemsetAsync(h->d_GF, 0, (size_t)GF_COUNT * Mp * sizeof(double), st));
      also(cudaMemcpyAsync(h->d_scen_gas, gas_tab.data(), gas_tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (h->d_trk_fail) also(cudaMemsetAsync(h->d_trk_fail, 0, Mp * sizeof(int32_t), st));
    /* the host vectors above must outlive the copies */
    also(cudaStreamSynchronize(st));
    if (e != cudaSuccess) {
      h->free_device();
      return fail(HX_ERR_CUDA, (std::string("hx_prepare uploads: ") + cudaGetErrorString(e)).c_str());
    }
  }

  HxDev &d = h->d;
  d.Mpad = Mpad; d.P = h->d_P; d.S = h->d_S; d.D = h->d_D; d.ker = h->d_ker; d.conv = h->d_conv;
  d.sst_hist = h->d_sst; d.tland_hist = h->d_tland; d.out = h->d_out; d.X = h->d_X; d.scen = h->d_scen;
  d.api_of_dev = h->d_api_of_dev;
  d.block_scen = h->d_block_scen; d.status = h->d_status; d.fail_year = h->d_fail_year;
  d.spinup_steps = h->d_spinup_steps; d.counters = h->d_counters; d.sched = h->d_sched;
  d.T = h->d_T; d.TK = h->d_TK; d.TO = h->d_TO; d.TOK = h->d_TOK; d.REC = h->d_REC; d.YCNT = h->d_YCNT;
  d.BP = h->d_BP; d.BF = h->d_BF; d.trk_fail = h->d_trk_fail;
  d.GP = h->d_GP; d.GF = h->d_GF; d.scen_gas = h->d_scen_gas;
  h->rec_elems = block_scen.size() * hx::track_record_bytes_per_cta() / sizeof(double);
  h->ycnt_bytes = block_scen.size() * hx::track_ycnt_bytes_per_tile();
  h->tables_constrained = any_constraint;
  h->tables_nbp = any_nbp;
  d.constrained = any_nbp ? 2 : any_constraint ? 1 : 0; /* refined in run_setup_and_spinup */
  for (int i = 0; i < HX_OUT_IDS; ++i) d.out_slot[i] = -1;
  for (int s = 0; s < nsel; ++s) d.out_slot[h->out_sel[s]] = s;
  d.n_out = nsel;
  d.slab_done = nullptr;
  d.out_minimal = 1;
  for (int s = 0; s < nsel; ++s)
    if (h->out_sel[s] == OUT_RF_TOT || h->out_sel[s] == OUT_RF_CO2) d.out_minimal = d.out_minimal ? 2 : 0;
    else if (h->out_sel[s] != OUT_CO2 && h->out_sel[s] != OUT_TAS) d.out_minimal = 0;

  h->prepared = true; /* upload_param / set_param paths need the buffers */
  for (int pi = 0; pi < PI_COUNT; ++pi) {
    int rc = h->upload_param(pi);
    if (rc) { h->prepared = false; return rc; }
  }
  if (h->d_BP) {
    int rc = h->upload_biomes();
    if (rc) { h->prepared = false; return rc; }
  }
  if (h->d_GP) {
    int rc = h->upload_gas();
    if (rc) { h->prepared = false; return rc; }
  }
  int rc = h->run_setup_and_spinup();
  if (rc) { h->prepared = false; return rc; }
  if (cudaStreamSynchronize(st) != cudaSuccess) {
    h->prepared = false;
    return fail(HX_ERR_CUDA, (std::string("set-up/spin-up kernels: ") +
                              cudaGetErrorString(cudaGetLastError())).c_str());
  }
  return HX_OK;
}

int hx_reset(hx_handle h) {
  NvtxRange nvtx_("hx_reset");
  if (!h) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_reset before hx_prepare");
  cudaSetDevice(h->cfg.device);
  if (h->params_dirty) {
    /* Core::reset(date < start) re-runs prepareToRun incl. spin-up (core.cpp:511-549) */
    return h->run_setup_and_spinup();
  }
  cudaError_t e = cudaMemcpyAsync(h->d_S, h->d_S_snap, (size_t)SI_COUNT * h->Mpad * sizeof(double),
                                  cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess && h->d_BF)
    e = cudaMemcpyAsync(h->d_BF, h->d_BF_snap,
                        (size_t)h->n_biomes * BF_COUNT * h->Mpad * sizeof(double),
                        cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess && h->d_GF)
    e = cudaMemcpyAsync(h->d_GF, h->d_GF_snap, (size_t)GF_COUNT * h->Mpad * sizeof(double),
                        cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h->d_status, h->d_status_post, (size_t)h->Mpad * sizeof(int32_t),
                        cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess && h->d_T) e = hx::launch_track_init(h->d, h->stream);
  if (e == cudaSuccess && h->d_trk_fail)
    e = cudaMemsetAsync(h->d_trk_fail, 0, (size_t)h->Mpad * sizeof(int32_t), h->stream);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  h->cur_row = 0;
  return HX_OK;
}

int hx_reset_date(hx_handle h, double date) {
  NvtxRange nvtx_("hx_reset_date");
  if (!h) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_reset_date before hx_prepare");
  const int y = (int)date;
  if (y <= h->cfg.start_year) return hx_reset(h);
  if (y > h->cfg.start_year + h->cur_row)
    return h->fail(HX_ERR_ARG, "reset date is after the current date"); /* core.cpp:520-523 */
  /* The run is deterministic (bit-reproducible), so the state at `date` is re-derived by running
   * to it again.  After an input change that is still exact as long as the change only touches
   * years after `date` -- R's setvar(core, dates, var, ...) followed by the reset to
   * min(dates) - 1 it asks for (R/messages.R:107-140): the years up to `date` see the inputs
   * they saw before.  A change that reaches further back (any undated parameter) would need
   * the recorded state of the OLD run, which the engine does not keep. */
  if (h->params_dirty && y - h->cfg.start_year >= h->dirty_from_row)
    return h->fail(HX_ERR_UNSUPPORTED,
                   "reset to a date inside the run after a change that also affects earlier "
                   "years: the engine keeps no per-year state history; reset to the start date");
  int rc = hx_reset(h);
  if (rc) return rc;
  return hx_run(h, (double)y);
}

int hx_run(hx_handle h, double run_to_date) {
  NvtxRange nvtx_("hx_run");
  if (!h) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_run before hx_prepare");
  cudaSetDevice(h->cfg.device);
  if (h->params_dirty) {
    /* The reference would carry the old state on with the new values; the engine holds no
     * history to do that from, and silently restarting from start_year would also overrun a
     * caller's buffer sized from hx_current_date.  Mid-run changes need an explicit reset. */
    if (h->cur_row > 0)
      return h->fail(HX_ERR_STATE, "parameters or inputs changed since the run began: call "
                                   "hx_reset (or hx_reset_date) before running on");
    int rc = h->run_setup_and_spinup();
    if (rc) return rc;
  }
  int to = run_to_date < 0 ? h->cfg.end_year : (int)run_to_date;
  if (to > h->cfg.end_year) return h->fail(HX_ERR_ARG, "run_to_date beyond end_year");
  const int r1 = to - h->cfg.start_year;
  if (r1 <= h->cur_row) return HX_OK; /* "Requested run-to date <= current date. Models not run." */
  cudaStream_t st = h->stream;
  cudaError_t e = cudaMemsetAsync(h->d_counters, 0, HX_NCOUNTERS * sizeof(unsigned long long), st);
  if (e == cudaSuccess) e = cudaEventRecord(h->ev0, st);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run: ") + cudaGetErrorString(e));
  e = h->launch_rows(h->cur_row, r1);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("run kernel launch: ") + cudaGetErrorString(e));
  if (!h->out_sel.empty()) {
    e = hx::launch_nan_fill(h->d, h->C, (int)h->out_sel.size(), h->cur_row, r1, st);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("nan fill launch: ") + cudaGetErrorString(e));
  }
  e = cudaEventRecord(h->ev1, st);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run: ") + cudaGetErrorString(e));
  h->cur_row = r1;
  return HX_OK;
}

/* Run and stream the results out while running: the run is cut into `segments` launches (whole
 * 16-year slabs) and the recorded years of each finished segment go to the host over the copy
 * engine while the next segment computes, so that of the device-to-host time only the last
 * segment's share is exposed.  Year-major output (the device layout) needs no transpose kernel,
 * which could not run next to the persistent run kernel anyway. */
int hx_run_stream(hx_handle h, double run_to_date, int32_t n_vars, const char *const *names,
                  double *const *outs, int32_t segments) {
  NvtxRange nvtx_("hx_run_stream");
  if (!h || n_vars <= 0 || !names || !outs) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_run_stream before hx_prepare");
  cudaSetDevice(h->cfg.device);
  if (h->params_dirty) {
    /* The reference would carry the old state on with the new values; the engine holds no
     * history to do that from, and silently restarting from start_year would also overrun a
     * caller's buffer sized from hx_current_date.  Mid-run changes need an explicit reset. */
    if (h->cur_row > 0)
      return h->fail(HX_ERR_STATE, "parameters or inputs changed since the run began: call "
                                   "hx_reset (or hx_reset_date) before running on");
    int rc = h->run_setup_and_spinup();
    if (rc) return rc;
  }
  const int to = run_to_date < 0 ? h->cfg.end_year : (int)run_to_date;
  if (to > h->cfg.end_year) return h->fail(HX_ERR_ARG, "run_to_date beyond end_year");
  const int r0 = h->cur_row, r1 = to - h->cfg.start_year;
  if (r1 <= r0) return HX_OK;
  std::vector<int> slot(n_vars);
  for (int v = 0; v < n_vars; ++v) {
    const int id = h->find_out(names[v]);
    if (id < 0 || h->d.out_slot[id] < 0)
      return h->fail(HX_ERR_ARG, std::string("output not recorded: ") + (names[v] ? names[v] : "?"));
    if (!outs[v]) return HX_ERR_ARG;
    slot[v] = h->d.out_slot[id];
  }
  {
    int rc = ensure_copy_stream(h);
    if (rc) return rc;
  }
  if (segments < 1) segments = 1;
  if (segments > 1 && !h->d_T) {
    /* ONE launch of the persistent run kernel; the slabs' output rows are copied out as the
     * kernel reports them complete (no per-segment launch tails, only the last slab's copy is
     * exposed).  Tracked runs keep the segmented form: their slabs are separate launches. */
    int rc = h->run_streamed(r0, r1, n_vars, slot.data(), outs);
    if (rc != HX_OK) return rc;
    h->cur_row = r1;
    return HX_OK;
  }
  const int nslab = (r1 - r0 + HX_SLAB_YEARS - 1) / HX_SLAB_YEARS;
  if (segments > nslab) segments = nslab;
  cudaStream_t st = h->stream;
  const int ny = r1 - r0; /* columns... rows of the caller's [year][member] blocks */
  {
    cudaError_t e0 = cudaMemsetAsync(h->d_counters, 0, HX_NCOUNTERS * sizeof(unsigned long long), st);
    if (e0 == cudaSuccess) e0 = cudaEventRecord(h->ev0, st);
    if (e0 != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run_stream: ") + cudaGetErrorString(e0));
  }
  /* Segment lengths halve: a segment's device-to-host copy hides behind the next segment's
   * computation (the copy engine is ~3x faster than the run per year), so only the LAST
   * segment's copy is exposed -- keep that one short instead of splitting evenly. */
  int ra = r0;
  int slabs_left = nslab;
  for (int sg = 0; sg < segments; ++sg) {
    const int take = (sg == segments - 1) ? slabs_left : (slabs_left + 1) / 2;
    slabs_left -= take;
    const int rb = (sg == segments - 1 || slabs_left == 0) ? r1 : ra + take * HX_SLAB_YEARS;
    if (rb <= ra) continue;
    cudaError_t e = h->launch_rows(ra, rb);
    if (e == cudaSuccess && !h->out_sel.empty())
      e = hx::launch_nan_fill(h->d, h->C, (int)h->out_sel.size(), ra, rb, st);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("run segment: ") + cudaGetErrorString(e));
    e = cudaEventRecord(h->ev_seg, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->copy_stream, h->ev_seg, 0);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run_stream: ") + cudaGetErrorString(e));
    for (int v = 0; v < n_vars; ++v) {
      /* years ra+1 .. rb are output rows ra .. rb-1 of the variable's [year][Mpad] block */
      const double *src = h->d_out + ((size_t)slot[v] * (h->nrow - 1) + ra) * h->Mpad;
      double *dst = outs[v] + (size_t)(ra - r0) * h->M;
      e = cudaMemcpy2DAsync(dst, (size_t)h->M * sizeof(double), src, (size_t)h->Mpad * sizeof(double),
                            (size_t)h->M * sizeof(double), (size_t)(rb - ra), cudaMemcpyDeviceToHost,
                            h->copy_stream);
      if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run_stream copy: ") + cudaGetErrorString(e));
    }
    ra = rb;
  }
  (void)ny;
  cudaError_t e = cudaEventRecord(h->ev1, st);
  /* the caller's stream must see the copies as done */
  if (e == cudaSuccess) e = cudaEventRecord(h->ev_copy, h->copy_stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, h->ev_copy, 0);
  h->cur_row = r1;
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->copy_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_run_stream: ") + cudaGetErrorString(e));
  return HX_OK;
}

/* ---- exchange of the recorded outputs between the GPUs of one node over peer memory ---- */
int hx_ipc_export(hx_handle h, void *handle64, int64_t *bytes) {
  if (!h || !handle64) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_ipc_export before hx_prepare");
  cudaSetDevice(h->cfg.device);
  cudaIpcMemHandle_t mh;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle(&mh, h->d_out);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  memcpy(handle64, &mh, 64);
  if (bytes) *bytes = (int64_t)h->out_sel.size() * (h->nrow - 1) * h->Mpad * (int64_t)sizeof(double);
  return HX_OK;
}

int hx_ipc_open(hx_handle h, int32_t n_peers, const void *handles, int32_t self_index) {
  if (!h || n_peers < 1 || !handles || self_index < 0 || self_index >= n_peers) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_ipc_open before hx_prepare");
  cudaSetDevice(h->cfg.device);
  hx_ipc_close(h);
  h->peer_out.assign(n_peers, nullptr);
  h->peer_self = self_index;
  for (int p = 0; p < n_peers; ++p) {
    if (p == self_index) { h->peer_out[p] = h->d_out; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char *)handles + (size_t)p * 64, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      hx_ipc_close(h);
      return h->fail(HX_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    h->peer_out[p] = (double *)ptr;
  }
  h->peer_stream.assign(n_peers, nullptr);
  for (int p = 0; p < n_peers; ++p)
    if (cudaStreamCreateWithFlags(&h->peer_stream[p], cudaStreamNonBlocking) != cudaSuccess) {
      hx_ipc_close(h);
      return h->fail(HX_ERR_CUDA, "hx_ipc_open: could not create the copy streams");
    }
  return ensure_copy_stream(h);
}

int hx_ipc_close(hx_handle h) {
  if (!h) return HX_OK;
  for (size_t p = 0; p < h->peer_out.size(); ++p)
    if ((int)p != h->peer_self && h->peer_out[p]) cudaIpcCloseMemHandle(h->peer_out[p]);
  for (cudaStream_t s : h->peer_stream)
    if (s) cudaStreamDestroy(s);
  h->peer_stream.clear();
  h->peer_out.clear();
  h->peer_self = -1;
  return HX_OK;
}

int hx_ipc_pull(hx_handle h, const char *name, int32_t year_a, int32_t year_b, double *dst_dev) {
  NvtxRange nvtx_("hx_ipc_pull");
  if (!h || !name || !dst_dev) return HX_ERR_ARG;
  if (h->peer_out.empty()) return h->fail(HX_ERR_STATE, "hx_ipc_pull before hx_ipc_open");
  const int id = h->find_out(name);
  if (id < 0 || h->d.out_slot[id] < 0) return h->fail(HX_ERR_ARG, std::string("output not recorded: ") + name);
  const int ra = year_a - h->cfg.start_year - 1, rb = year_b - h->cfg.start_year; /* rows [ra, rb) */
  if (ra < 0 || rb > h->nrow - 1 || rb <= ra) return h->fail(HX_ERR_ARG, "hx_ipc_pull: bad year range");
  cudaSetDevice(h->cfg.device);
  const size_t block = (size_t)(h->nrow - 1) * h->Mpad;
  const size_t off = ((size_t)h->d.out_slot[id] * (h->nrow - 1) + ra) * h->Mpad;
  const int n = (int)h->peer_out.size();
  for (int k = 0; k < n; ++k) {
    const int p = (h->peer_self + 1 + k) % n; /* every rank starts with a different peer */
    cudaError_t e = cudaMemcpyAsync(dst_dev + (size_t)p * block + (size_t)ra * h->Mpad,
                                    h->peer_out[p] + off, (size_t)(rb - ra) * h->Mpad * sizeof(double),
                                    cudaMemcpyDefault, h->peer_stream[p]);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_ipc_pull: ") + cudaGetErrorString(e));
  }
  return HX_OK;
}

int hx_ipc_wait(hx_handle h) {
  if (!h) return HX_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  for (cudaStream_t s : h->peer_stream) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_ipc_wait: ") + cudaGetErrorString(e));
  }
  return HX_OK;
}

/* ---- push exchange: every rank's finished slabs are written into every peer's gather block ---- */
int hx_xchg_close(hx_handle h) {
  if (!h) return HX_OK;
  for (size_t p = 0; p < h->peer_gather.size(); ++p)
    if ((int)p != h->xchg_self && h->peer_gather[p]) cudaIpcCloseMemHandle(h->peer_gather[p]);
  for (cudaStream_t s : h->push_stream)
    if (s) cudaStreamDestroy(s);
  h->push_stream.clear();
  h->peer_gather.clear();
  if (h->d_gather) cudaFree(h->d_gather);
  h->d_gather = nullptr;
  h->xchg_self = -1;
  return HX_OK;
}

int hx_xchg_create(hx_handle h, int32_t n_peers, int32_t self_index, void *handle64, int64_t *bytes) {
  if (!h || n_peers < 1 || self_index < 0 || self_index >= n_peers || !handle64) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_xchg_create before hx_prepare");
  if (h->out_sel.empty()) return h->fail(HX_ERR_STATE, "hx_xchg_create: no outputs are recorded");
  cudaSetDevice(h->cfg.device);
  hx_xchg_close(h);
  h->gather_elems_per_rank = h->out_sel.size() * (size_t)(h->nrow - 1) * h->Mpad;
  const size_t total = (size_t)n_peers * h->gather_elems_per_rank * sizeof(double);
  cudaError_t e = cudaMalloc(&h->d_gather, total);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_xchg_create: ") + cudaGetErrorString(e));
  cudaIpcMemHandle_t mh;
  e = cudaIpcGetMemHandle(&mh, h->d_gather);
  if (e != cudaSuccess) {
    hx_xchg_close(h);
    return h->fail(HX_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &mh, 64);
  if (bytes) *bytes = (int64_t)total;
  h->xchg_self = self_index;
  h->peer_gather.assign(n_peers, nullptr);
  h->peer_gather[self_index] = h->d_gather;
  return HX_OK;
}

int hx_xchg_open(hx_handle h, int32_t n_peers, const void *handles) {
  if (!h || !handles) return HX_ERR_ARG;
  if (h->xchg_self < 0 || n_peers != (int)h->peer_gather.size())
    return h->fail(HX_ERR_STATE, "hx_xchg_open: call hx_xchg_create with the same n_peers first");
  cudaSetDevice(h->cfg.device);
  for (int p = 0; p < n_peers; ++p) {
    if (p == h->xchg_self) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char *)handles + (size_t)p * 64, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      hx_xchg_close(h);
      return h->fail(HX_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    h->peer_gather[p] = (double *)ptr;
  }
  h->push_stream.assign(n_peers, nullptr);
  for (int p = 0; p < n_peers; ++p)
    if (cudaStreamCreateWithFlags(&h->push_stream[p], cudaStreamNonBlocking) != cudaSuccess) {
      hx_xchg_close(h);
      return h->fail(HX_ERR_CUDA, "hx_xchg_open: could not create the copy streams");
    }
  return HX_OK;
}

int hx_xchg_block(hx_handle h, const double **dev_ptr, int64_t *elems_per_rank) {
  if (!h || !dev_ptr) return HX_ERR_ARG;
  if (!h->d_gather) return h->fail(HX_ERR_STATE, "hx_xchg_block before hx_xchg_create");
  *dev_ptr = h->d_gather;
  if (elems_per_rank) *elems_per_rank = (int64_t)h->gather_elems_per_rank;
  return HX_OK;
}

int hx_run_exchange(hx_handle h, double run_to_date) {
  NvtxRange nvtx_("hx_run_exchange");
  if (!h) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_run_exchange before hx_prepare");
  if (h->push_stream.empty()) return h->fail(HX_ERR_STATE, "hx_run_exchange before hx_xchg_open");
  if (h->d_T) return h->fail(HX_ERR_UNSUPPORTED, "hx_run_exchange with carbon tracking: use hx_run, then a gather");
  cudaSetDevice(h->cfg.device);
  if (h->params_dirty) {
    if (h->cur_row > 0)
      return h->fail(HX_ERR_STATE, "parameters or inputs changed since the run began: call "
                                   "hx_reset (or hx_reset_date) before running on");
    int rc = h->run_setup_and_spinup();
    if (rc) return rc;
  }
  const int to = run_to_date < 0 ? h->cfg.end_year : (int)run_to_date;
  if (to > h->cfg.end_year) return h->fail(HX_ERR_ARG, "run_to_date beyond end_year");
  const int r0 = h->cur_row, r1 = to - h->cfg.start_year;
  if (r1 <= r0) return HX_OK;
  int rc = h->run_pushed(r0, r1);
  if (rc != HX_OK) return rc;
  h->cur_row = r1;
  return HX_OK;
}

int hx_event_record(hx_handle h, int32_t idx) {
  if (!h || idx < 0 || idx >= 16) return HX_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  if (!h->ev_user[idx] &&
      cudaEventCreateWithFlags(&h->ev_user[idx], cudaEventDisableTiming) != cudaSuccess)
    return h->fail(HX_ERR_CUDA, "cudaEventCreate");
  cudaError_t e = cudaEventRecord(h->ev_user[idx], h->stream);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  return HX_OK;
}

int hx_event_synchronize(hx_handle h, int32_t idx) {
  if (!h || idx < 0 || idx >= 16 || !h->ev_user[idx]) return HX_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  cudaError_t e = cudaEventSynchronize(h->ev_user[idx]);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  return HX_OK;
}

int hx_synchronize(hx_handle h) {
  if (!h) return HX_ERR_ARG;
  cudaSetDevice(h->cfg.device);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  return HX_OK;
}

double hx_last_run_ms(hx_handle h) {
  if (!h || !h->prepared) return -1.0;
  cudaSetDevice(h->cfg.device);
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

double hx_current_date(hx_handle h) { return h ? h->cfg.start_year + h->cur_row : -1.0; }

} /* extern "C" */

/* What the reference answers for the START date (R's fetchvars keeps dates >= startdate,
 * R/messages.R:66; tests/testthat/test_parameters.R:46 asks for the CO2 concentration of 1745):
 * every component records its state at the start date once the spin-up is done -- the pools as
 * the spin-up left them (simpleNbox.cpp:640-662 through record_state, ocean_component.cpp
 * getData), the preindustrial concentrations (C0, M0, N0, PO3), zero for temperatures, heat
 * fluxes, forcings, pH / pCO2 and the ocean uptake, the spin-up's last NPP and RH; NBP has no
 * entry there (the reference throws "Interpolation requested but not allowed").  Pinned by
 * tests/golden/ref_startdate.npz. */
int Engine::start_row(int id, std::vector<double> &row) {
  row.assign((size_t)M, 0.0);
  std::vector<double> tmp((size_t)M);
  int rc = HX_OK;
  auto state = [&](int f) { return gather_field_host(d_S_snap, f, SI_COUNT, row.data()); };
  if (id >= OUT_COUNT) { /* <biome>.<name> */
    static const int bf[BO_COUNT] = {BF_VEG, BF_DET, BF_SOIL, BF_PERMAFROST, BF_THAWED, BF_X_NPP, BF_X_RH};
    const int ib = (id - OUT_COUNT) / BO_COUNT, k = (id - OUT_COUNT) % BO_COUNT;
    rc = gather_field_host(d_BF_snap, ib * BF_COUNT + bf[k], n_biomes * BF_COUNT, row.data());
  } else switch (id) {
    case OUT_CO2: rc = param_values(PI_C0, row.data()); break;
    case OUT_CH4: rc = param_values(PI_M0, row.data()); break;
    case OUT_N2O: rc = param_values(PI_N0, row.data()); break;
    case OUT_O3: rc = param_values(PI_PO3, row.data()); break;
    case OUT_ATMOS_C: rc = state(SI_ATMOS); break;
    case OUT_VEG_C: rc = state(SI_VEG); break;
    case OUT_DETRITUS_C: rc = state(SI_DET); break;
    case OUT_SOIL_C: rc = state(SI_SOIL); break;
    case OUT_PERMAFROST_C: rc = state(SI_PERMAFROST); break;
    case OUT_THAWEDP_C: rc = state(SI_THAWED); break;
    case OUT_EARTH_C: rc = state(SI_EARTH); break;
    case OUT_CARBON_HL: rc = state(SI_BOX_HL); break;
    case OUT_CARBON_LL: rc = state(SI_BOX_LL); break;
    case OUT_CARBON_IO: rc = state(SI_BOX_IO); break;
    case OUT_CARBON_DO: rc = state(SI_BOX_DO); break;
    case OUT_NPP: rc = state(SI_X_NPP); break;
    case OUT_RH: rc = state(SI_X_RH); break;
    case OUT_OCEAN_C: { /* the run kernel's order of summation */
      static const int box[4] = {SI_BOX_DO, SI_BOX_IO, SI_BOX_LL, SI_BOX_HL};
      for (int b = 0; b < 4 && rc == HX_OK; ++b) {
        rc = gather_field_host(d_S_snap, box[b], SI_COUNT, tmp.data());
        for (int i = 0; i < M; ++i) row[i] = b ? row[i] + tmp[i] : tmp[i];
      }
      break;
    }
    case OUT_NBP:
      return fail(HX_ERR_ARG, "NBP has no value at the start date (the reference: Interpolation "
                              "requested but not allowed)");
    default: break; /* temperatures, heat fluxes, forcings, pH, pCO2, uptake, rh_ch4: 0 */
  }
  if (rc) return rc;
  /* a member whose spin-up failed has no start state */
  std::vector<int32_t> st((size_t)Mpad);
  CUDA_TRY(cudaMemcpy(st.data(), d_status_snap, (size_t)Mpad * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < M; ++i)
    if (st[dev_of_api[i]] > 0) row[i] = std::numeric_limits<double>::quiet_NaN();
  return HX_OK;
}

/* out[member][k] = recorded output `slot` of year index yidx[k], API member order */
int Engine::fetch_rows(int slot, const std::vector<int32_t> &yidx, int n_dates, double *out) {
  int rc = ensure_stage((size_t)M * n_dates * sizeof(double));
  if (rc) return rc;
  if ((size_t)n_dates > yidx_cap) {
    if (d_yidx) cudaFree(d_yidx);
    d_yidx = nullptr;
    yidx_cap = 0;
    CUDA_TRY(cudaMalloc(&d_yidx, (size_t)n_dates * sizeof(int32_t)));
    yidx_cap = n_dates;
  }
  CUDA_TRY(cudaMemcpyAsync(d_yidx, yidx.data(), (size_t)n_dates * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  const double *src = d_out + (size_t)slot * (nrow - 1) * Mpad;
  dim3 grid((M + 31) / 32, (n_dates + 31) / 32), block(32, 8);
  k_fetch_transpose<<<grid, block, 0, stream>>>(d_stage, src, d_yidx, nullptr, n_dates, M, (size_t)Mpad);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, d_stage, (size_t)M * n_dates * sizeof(double), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return HX_OK;
}

extern "C" {
int hx_fetch(hx_handle h, const char *name, const double *dates, int32_t n_dates, double *out) {
  NvtxRange nvtx_("hx_fetch");
  if (!h || !name || !dates || !out || n_dates <= 0) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_fetch before hx_prepare");
  {
    int rc = HX_OK;
    if (h->fetch_derived(name, dates, n_dates, out, rc)) return rc;
    if (h->fetch_functions(name, dates, n_dates, out, rc)) return rc;
  }
  {
    /* an INPUT series read back -- getData of the component that owns it (simpleNbox.cpp:650-662
     * ffi / daccs / luc, ch4_component.cpp, n2o_component.cpp, oh_component.cpp, bc / oc / so2 /
     * nh3, halocarbon_component.cpp "<gas>_emissions"): tests/testthat/test_pulse.R fetches
     * luc_emissions.  Any year of the run, start date included; a constraint series answers NaN
     * (the reference's MISSING_FLOAT) where it has no entry. */
    const int si = Engine::find_raw(name);
    const int ci = si < 0 ? Engine::find_constraint(name) : -1;
    if (si >= 0 || ci >= 0) {
      for (int k = 0; k < n_dates; ++k) {
        const int r = (int)dates[k] - h->cfg.start_year;
        if (r < 0 || r >= h->nrow) return h->fail(HX_ERR_ARG, "date outside [start_year, end_year]");
        for (int i = 0; i < h->M; ++i) {
          const int sc = h->member_scen[i];
          out[(size_t)i * n_dates + k] = si >= 0 ? h->raw[sc][(size_t)si * h->nrow + r]
                                                 : h->cons[sc][(size_t)ci * h->nrow + r];
        }
      }
      return HX_OK;
    }
  }
  const int id = h->find_out(name);
  if (id < 0) return h->fail(HX_ERR_ARG, std::string("unknown output variable: ") + name);
  const int slot = h->d.out_slot[id];
  if (slot < 0) return h->fail(HX_ERR_ARG, std::string(name) + " was not selected with hx_select_outputs");
  cudaSetDevice(h->cfg.device);
  std::vector<int32_t> yidx(n_dates);
  bool want_start = false;
  for (int k = 0; k < n_dates; ++k) {
    const int r = (int)dates[k] - h->cfg.start_year;
    if (r < 0 || r > h->cur_row)
      return h->fail(HX_ERR_ARG, "date outside [start_year, current date]");
    want_start = want_start || r == 0;
    yidx[k] = r > 0 ? r - 1 : 0; /* the start date's column is filled in below */
  }
  std::vector<double> row0;
  if (want_start) {
    const int rc0 = h->start_row(id, row0);
    if (rc0) return rc0;
  }
  if (h->cur_row > 0) { /* otherwise nothing but the start date can have been asked for */
    const int rc = h->fetch_rows(slot, yidx, n_dates, out);
    if (rc) return rc;
  }
  if (want_start)
    for (int k = 0; k < n_dates; ++k)
      if ((int)dates[k] == h->cfg.start_year)
        for (int i = 0; i < h->M; ++i) out[(size_t)i * n_dates + k] = row0[i];
  return HX_OK;
}

int hx_set_tracking(hx_handle h, int32_t tracking_date, int32_t record_every) {
  if (!h) return HX_ERR_ARG;
  if (h->prepared) return h->fail(HX_ERR_STATE, "tracking must be configured before hx_prepare");
  if (record_every < 0) return h->fail(HX_ERR_ARG, "record_every must be >= 0");
  h->tracking_date = tracking_date;
  h->track_every = record_every;
  return HX_OK;
}

int hx_tracking_years(hx_handle h, int32_t *years, int32_t cap) {
  if (!h) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_tracking_years before hx_prepare");
  const int n = (int)h->track_years.size();
  for (int i = 0; i < n && i < cap && years; ++i) years[i] = h->track_years[i];
  return n;
}

int hx_fetch_tracking(hx_handle h, double date, double *frac, uint32_t *mask) {
  NvtxRange nvtx_("hx_fetch_tracking");
  if (!h || !frac) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_fetch_tracking before hx_prepare");
  if (!h->d_T) return h->fail(HX_ERR_STATE, "carbon tracking is off (no trackingDate set)");
  const int y = (int)date;
  const int cur = h->cfg.start_year + h->cur_row;
  if (y < h->tracking_date || y > cur)
    return h->fail(HX_ERR_ARG, "date outside [trackingDate, current date]");
  cudaSetDevice(h->cfg.device);
  const int nq = HX_NPOOL * HX_NSRC;
  const size_t fbytes = (size_t)h->M * nq * sizeof(double);
  const size_t mbytes = (size_t)h->M * HX_NPOOL * sizeof(uint32_t);
  int rc = h->ensure_stage(fbytes + mbytes);
  if (rc) return rc;
  double *sf = h->d_stage;
  uint32_t *sm = (uint32_t *)((char *)h->d_stage + fbytes);
  cudaStream_t st = h->stream;
  cudaError_t e = cudaSuccess;
  int rec = -1;
  for (size_t i = 0; i < h->track_years.size(); ++i)
    if (h->track_years[i] == y) rec = (int)i;
  if (rec >= 0) {
    std::vector<int32_t> yidx(nq);
    for (int k = 0; k < nq; ++k) yidx[k] = k;
    if ((size_t)nq > h->yidx_cap) {
      if (h->d_yidx) cudaFree(h->d_yidx);
      h->d_yidx = nullptr;
      if (cudaMalloc(&h->d_yidx, (size_t)nq * sizeof(int32_t)) != cudaSuccess)
        return h->fail(HX_ERR_CUDA, "cudaMalloc yidx");
      h->yidx_cap = nq;
    }
    e = cudaMemcpyAsync(h->d_yidx, yidx.data(), (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_fetch_tracking: ") + cudaGetErrorString(e));
    dim3 grid((h->M + 31) / 32, (nq + 31) / 32), block(32, 8);
    k_fetch_transpose<<<grid, block, 0, st>>>(sf, h->d_TO + (size_t)rec * nq * h->Mpad, h->d_yidx,
                                             h->d_dev_of_api, nq, h->M, (size_t)h->Mpad);
    k_fetch_masks<<<(h->M * HX_NPOOL + 255) / 256, 256, 0, st>>>(
        sm, h->d_TOK + (size_t)rec * HX_NPOOL * h->Mpad, h->d_dev_of_api, HX_NPOOL, h->M,
        (size_t)h->Mpad);
    if (e == cudaSuccess) e = cudaGetLastError();
  } else if (y == cur) {
    /* not a recorded year, but the live maps are this year's */
    k_gather_track<<<(h->M * nq + 255) / 256, 256, 0, st>>>(sf, sm, h->d_T, h->d_TK,
                                                           h->d_dev_of_api, h->M);
    e = cudaGetLastError();
  } else {
    return h->fail(HX_ERR_ARG, "year was not recorded (see hx_set_tracking record_every)");
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(frac, sf, fbytes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && mask) e = cudaMemcpyAsync(mask, sm, mbytes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, std::string("hx_fetch_tracking: ") + cudaGetErrorString(e));
  /* failed members report NaN like every other output */
  std::vector<int32_t> stt(h->M), fy(h->M);
  rc = hx_member_status(h, stt.data(), fy.data(), h->M);
  if (rc) return rc;
  for (int i = 0; i < h->M; ++i)
    if (stt[i] > 0 && fy[i] <= y) {
      for (int k = 0; k < nq; ++k) frac[(size_t)i * nq + k] = NAN;
      if (mask) for (int k = 0; k < HX_NPOOL; ++k) mask[(size_t)i * HX_NPOOL + k] = 0;
    }
  return HX_OK;
}

int hx_output_device(hx_handle h, const char *name, const double **dev_ptr, int64_t *member_stride,
                     int32_t *n_years) {
  if (!h || !name || !dev_ptr) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_output_device before hx_prepare");
  const int id = h->find_out(name);
  if (id < 0 || h->d.out_slot[id] < 0) return h->fail(HX_ERR_ARG, std::string("output not recorded: ") + name);
  *dev_ptr = h->d_out + (size_t)h->d.out_slot[id] * (h->nrow - 1) * h->Mpad;
  if (member_stride) *member_stride = h->Mpad;
  if (n_years) *n_years = h->nrow - 1;
  return HX_OK;
}

int hx_member_status(hx_handle h, int32_t *status, int32_t *fail_year, int32_t n) {
  if (!h || n != h->M) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_member_status before hx_prepare");
  cudaSetDevice(h->cfg.device);
  std::vector<int32_t> st(h->Mpad), fy(h->Mpad);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaMemcpy(st.data(), h->d_status, (size_t)h->Mpad * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess)
    e = cudaMemcpy(fy.data(), h->d_fail_year, (size_t)h->Mpad * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  for (int i = 0; i < n; ++i) {
    if (status) status[i] = st[h->dev_of_api[i]];
    if (fail_year) fail_year[i] = fy[h->dev_of_api[i]];
  }
  return HX_OK;
}

int hx_counters(hx_handle h, uint64_t *out, int32_t n) {
  if (!h || !out || n < 1) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_counters before hx_prepare");
  cudaSetDevice(h->cfg.device);
  unsigned long long tmp[HX_NCOUNTERS];
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaMemcpy(tmp, h->d_counters, sizeof tmp, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  for (int i = 0; i < n && i < HX_NCOUNTERS; ++i) out[i] = tmp[i];
  return HX_OK;
}

int hx_spinup_state(hx_handle h, int32_t member, double *out14) {
  if (!h || !out14 || member < 0 || member >= h->M) return HX_ERR_ARG;
  if (!h->prepared) return h->fail(HX_ERR_STATE, "hx_spinup_state before hx_prepare");
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  const int dm = h->dev_of_api[member];
  static const int idx[13] = {SI_ATMOS, SI_VEG, SI_DET, SI_SOIL, SI_PERMAFROST, SI_THAWED, SI_EARTH,
                              SI_BOX_HL, SI_BOX_LL, SI_BOX_IO, SI_BOX_DO, SI_ALK_HL, SI_ALK_LL};
  for (int k = 0; k < 13; ++k) {
    cudaError_t e = cudaMemcpy(out14 + k, h->d_S_snap + HX_TILED(idx[k], dm, SI_COUNT),
                               sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return h->fail(HX_ERR_CUDA, cudaGetErrorString(e));
  }
  int32_t steps = 0;
  if (cudaMemcpy(&steps, h->d_spinup_steps + dm, sizeof steps, cudaMemcpyDeviceToHost) != cudaSuccess)
    return h->fail(HX_ERR_CUDA, "hx_spinup_state: copy failed");
  out14[13] = steps;
  return HX_OK;
}

} /* extern "C" */
