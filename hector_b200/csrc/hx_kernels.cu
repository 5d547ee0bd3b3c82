/* hx_kernels.cu -- sm_100a kernels of the Hector ensemble engine.
 *
 *   hx_setup_kernel   per member: derived constants (ocean exchange rates, DOECLIM matrices and
 *                     lag kernel) and the pre-spin-up state          [prepareToRun of every component]
 *   hx_spinup_kernel  per member: spin the carbon cycle up to steady state, then equilibrate the
 *                     surface-box alkalinities (Brent)               [Core::run_spinup + chem_equilibrate]
 *   hx_run_kernel     per member: the yearly coupled step for a run segment; the scenario table
 *                     is streamed through shared memory in slabs with cp.async.bulk + mbarrier
 *                                                                   [Core::run year loop]
 *
 * One thread owns one member; all per-member arrays are SoA so warps read/write 256 B rows.
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "hx_kernels.h"
#include "hx_model.cuh"

namespace hx {

/* The DOECLIM convolution accumulates with fused multiply-adds.  HX_NO_FMA (together with nvcc
 * --fmad=false) gives a build whose arithmetic rounds like the reference's x86-64 build, product
 * and sum separately -- a debugging discriminator for parity work (tools/README.md), not a
 * product configuration. */
#ifdef HX_NO_FMA
#define HX_CONV_FMA(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#else
#define HX_CONV_FMA(a, b, c) fma((a), (b), (c))
#endif

#define HX_MAX_DEVICES 64 /* device ordinals with cached launch parameters */

/* run-kernel dynamic shared memory map (bytes) */
#define HX_SMEM_SLAB_BYTES ((HX_SLAB_YEARS + 1) * SC_STRIDE * 8)
#define HX_SMEM_ROW0 HX_SMEM_SLAB_BYTES
#define HX_SMEM_CHEMK (HX_SMEM_ROW0 + SC_STRIDE * 8)
#define HX_SMEM_RK (HX_SMEM_CHEMK + 10 * HX_BLOCK * 8)
#define HX_SMEM_RUN_BYTES (HX_SMEM_RK + HX_RK_SLOTS * HX_BLOCK * 8)
/* builds with two resident CTAs per SM have room for the tile's whole state block S next to
 * that (2 x 111 KB of the SM's 228 KB): the year loop then reads and writes its state in shared
 * memory, and HBM / L2 see it once per 16-year work item */
#ifndef HX_SMEM_STATE
#define HX_SMEM_STATE 1
#endif
#define HX_SMEM_STATE_BYTES ((HX_HOT_COUNT + SI_COUNT - SI_REG_COUNT) * HX_BLOCK * 8)
template <int MINCTAS>
__host__ __device__ constexpr bool smem_state() { return HX_SMEM_STATE && MINCTAS == 2; }
/* the latency build (one CTA per SM at most) keeps ALL of P | D on chip, not just the hot stretch */
#define HX_SMEM_LAT_BYTES ((PD_COUNT + SI_COUNT - SI_REG_COUNT) * HX_BLOCK * 8)
template <int MINCTAS, bool LAT = false>
__host__ __device__ constexpr size_t run_smem_bytes() {
  return HX_SMEM_RUN_BYTES + (!smem_state<MINCTAS>() ? 0 : LAT ? HX_SMEM_LAT_BYTES : HX_SMEM_STATE_BYTES);
}

/* ---- small PTX wrappers: mbarrier + bulk async copy (TMA engine, UBLKCP in SASS) ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
/* non-aligned CTA barrier on named barrier 1 (barrier 0 is __syncthreads'): may be reached from
 * divergent code, each of the HX_BLOCK threads arrives once per phase */
__device__ __forceinline__ void year_barrier() {
  asm volatile("barrier.sync 1, %0;" ::"n"(HX_BLOCK) : "memory");
}
/* a read-only load the compiler leaves where it is written */
__device__ __forceinline__ double ldg_pinned(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
/* a member constant requested ahead of its use: pinned where it is written when it comes from
 * L2, an ordinary read when the constants are in shared memory */
#define PAR_AHEAD(ptr) (BS.psm ? *(ptr) : ldg_pinned(ptr))
/* the same for data this kernel writes itself (the histories): a coherent load */
__device__ __forceinline__ double ld_pinned(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* CTA-tiled SoA accessors: one base pointer per array, compile-time field offsets */
struct Bases {
  const double *P;
  const double *H; /* the hot stretch of P | D: in the array itself or the run kernel's shared copy */
  double *Sg;      /* the state in global memory (S may point into shared memory) */
  bool psm;        /* P and D point into shared memory (the latency build): plain loads only */
  double *S, *D, *ker, *sst, *tland, *conv;
  const double *BP; /* per-biome parameters / state of this member (null: single biome) */
  double *BF;
  const double *GP; /* per-member N2O / halocarbon parameters and state (null: host series) */
  double *GF;
};
__device__ __forceinline__ Bases make_bases(const HxDev &d, const HxConst &C, int m) {
  const size_t tile = (size_t)(m / HX_BLOCK), ln = (size_t)(m % HX_BLOCK);
  Bases b;
  b.P = d.P + tile * PD_COUNT * HX_BLOCK + ln;
  b.S = d.S + tile * SI_COUNT * HX_BLOCK + ln;
  b.D = const_cast<double *>(b.P) + PI_COUNT * HX_BLOCK; /* same tile block: a compile-time offset */
  b.H = b.P + HX_HOT_FIRST * HX_BLOCK;
  b.Sg = b.S;
  b.psm = false;
  b.ker = d.ker + tile * (size_t)HX_KER_ROWS(C.nrow) * HX_BLOCK + ln;
  b.conv = d.conv + tile * (size_t)HX_SLAB_YEARS * HX_BLOCK + ln;
  b.sst = d.sst_hist + tile * (size_t)C.nrow * HX_BLOCK + ln;
  b.tland = d.tland_hist + tile * (size_t)C.nrow * HX_BLOCK + ln;
  b.BP = d.BP ? d.BP + tile * (size_t)C.n_biomes * BP_COUNT * HX_BLOCK + ln : nullptr;
  b.BF = d.BF ? d.BF + tile * (size_t)C.n_biomes * BF_COUNT * HX_BLOCK + ln : nullptr;
  b.GP = d.GP ? d.GP + tile * (size_t)GP_COUNT * HX_BLOCK + ln : nullptr;
  b.GF = d.GF ? d.GF + tile * (size_t)GF_COUNT * HX_BLOCK + ln : nullptr;
  return b;
}
#define PAR(i) (BS.psm ? BS.P[(i) * HX_BLOCK] : __ldg(BS.P + (i) * HX_BLOCK))
#define STATE(i) BS.S[(i) * HX_BLOCK]
#define DER(i) BS.D[(i) * HX_BLOCK]

__device__ __forceinline__ LandPar load_landpar(const Bases &BS) {
  LandPar p;
  p.P = BS.P; p.D = BS.D; p.H = BS.H; p.psm = BS.psm;
  return p;
}

__device__ __forceinline__ void load_member(const Bases &BS, Member &mb) {
  mb.S = BS.S;
  /* the SI_REG_COUNT register-resident fields: straight from / to global memory */
#define STATE_G(i) BS.Sg[(i) * HX_BLOCK]
  mb.atmos = STATE_G(SI_ATMOS); mb.veg = STATE_G(SI_VEG); mb.det = STATE_G(SI_DET);
  mb.soil = STATE_G(SI_SOIL); mb.perm = STATE_G(SI_PERMAFROST); mb.thawed = STATE_G(SI_THAWED);
  mb.earth = STATE_G(SI_EARTH);
  mb.bHL = STATE_G(SI_BOX_HL); mb.bLL = STATE_G(SI_BOX_LL); mb.bIO = STATE_G(SI_BOX_IO);
  mb.bDO = STATE_G(SI_BOX_DO);
  mb.max_timestep = STATE_G(SI_MAX_TIMESTEP); mb.timeout = (int)STATE_G(SI_TIMEOUT);
  mb.solver_dt = STATE_G(SI_SOLVER_DT);
  mb.status = 0; mb.neg = false; mb.timesteps = 0;
  mb.pco2HL = mb.pco2LL = 0.0; mb.gHL = mb.gLL = 0.0; mb.luc_e = mb.luc_u = 0.0;
  mb.X = nullptr; mb.REC = nullptr; mb.rec_stride = 0; mb.rec_n = 0; mb.trk = false; mb.trk_bad = false;
  mb.BIOP = BS.BP; mb.BIOF = BS.BF;
}

__device__ __forceinline__ void store_member(const Bases &BS, const Member &mb) {
  STATE_G(SI_ATMOS) = mb.atmos; STATE_G(SI_VEG) = mb.veg; STATE_G(SI_DET) = mb.det;
  STATE_G(SI_SOIL) = mb.soil; STATE_G(SI_PERMAFROST) = mb.perm; STATE_G(SI_THAWED) = mb.thawed;
  STATE_G(SI_EARTH) = mb.earth;
  STATE_G(SI_BOX_HL) = mb.bHL; STATE_G(SI_BOX_LL) = mb.bLL; STATE_G(SI_BOX_IO) = mb.bIO;
  STATE_G(SI_BOX_DO) = mb.bDO;
  STATE_G(SI_MAX_TIMESTEP) = mb.max_timestep; STATE_G(SI_TIMEOUT) = (double)mb.timeout;
  STATE_G(SI_SOLVER_DT) = mb.solver_dt;
}

__device__ __forceinline__ void flush_work(const HxDev &d, const Work &w, unsigned years,
                                           unsigned failed) {
  /* integer atomics only: sums are order-independent, results stay bit-reproducible */
  unsigned long long *c = d.counters;
  atomicAdd(c + HX_CNT_RHS_EVALS, (unsigned long long)w.rhs);
  atomicAdd(c + HX_CNT_RK_STEPS, (unsigned long long)w.steps);
  atomicAdd(c + HX_CNT_RK_REJECTED, (unsigned long long)w.rejected);
  atomicAdd(c + HX_CNT_STASHES, (unsigned long long)w.stashes);
  atomicAdd(c + HX_CNT_NEWTON_ITERS, (unsigned long long)w.newton_it);
  atomicAdd(c + HX_CNT_NEWTON_CALLS, (unsigned long long)w.newton_calls);
  atomicAdd(c + HX_CNT_FAILED_MEMBERS, (unsigned long long)failed);
  atomicAdd(c + HX_CNT_MEMBER_YEARS, (unsigned long long)years);
}

/* ======================================================================================== */
/* set-up: OceanComponent::prepareToRun (ocean_component.cpp:202-319),
 * TemperatureComponent::prepareToRun (temperature_component.cpp:196-413),
 * SimpleNbox::prepareToRun (simpleNbox-runtime.cpp:61-197) and the CH4/solver initial values. */
__device__ __forceinline__ void setup_doeclim(const Bases &BS, const HxConst &C);
__device__ __forceinline__ void setup_state(const HxDev &d, const HxConst &C, const Bases &BS, int m);

/* phase bit 1: ocean rates, initial pools and state -- all the spin-up needs; bit 2: the DOECLIM
 * matrices and lag kernel (the expensive part: nrow kernel entries per member), which the
 * engine runs next to the spin-up on a second stream */
__global__ void __launch_bounds__(HX_BLOCK)
hx_setup_kernel(const __grid_constant__ HxDev d, const __grid_constant__ HxConst C, int phase) {
  const int m = blockIdx.x * HX_BLOCK + threadIdx.x;
  if (m >= d.Mpad) return;
  if (d.status[m] < 0) return; /* padding lane */
  const Bases BS = make_bases(d, C, m);

  if (phase & 2) {
    setup_doeclim(BS, C);
    if (!(phase & 1)) return;
  }
  /* ocean exchange rates (fraction of the box per year), ocean_component.cpp:262-284 */
  const double spy = C.spy_ocean;
  const double tt = PAR(PI_TT), tu = PAR(PI_TU), twi = PAR(PI_TWI), tid = PAR(PI_TID);
  const double LL_HL = (tt * spy) / C.vol_LL;
  const double HL_DO = ((tt + tu) * spy) / C.vol_HL;
  const double DO_IO = ((tt + tu) * spy) / C.vol_DO;
  const double IO_HL = (tu * spy) / C.vol_IO;
  const double IO_LL = (tt * spy) / C.vol_IO;
  const double IO_LLex = (twi * spy) / C.vol_IO;
  const double LL_IOex = (twi * spy) / C.vol_LL;
  const double DO_IOex = (tid * spy) / C.vol_DO;
  const double IO_DOex = (tid * spy) / C.vol_IO;
  DER(DI_K_LL_HL) = LL_HL; DER(DI_K_LL_IO) = LL_IOex; DER(DI_K_HL_DO) = HL_DO;
  DER(DI_K_IO_LL) = IO_LL + IO_LLex; DER(DI_K_IO_HL) = IO_HL; DER(DI_K_IO_DO) = IO_DOex;
  DER(DI_K_DO_IO) = DO_IO + DO_IOex;
  setup_state(d, C, BS, m);
}

/* DOECLIM, temperature_component.cpp:248-412 */
__device__ __forceinline__ void setup_doeclim(const Bases &BS, const HxConst &C) {
  const double dt = 1.0, ak = DC_AK, bk = DC_BK, csw = DC_CSW, rlam = DC_RLAM, bsi = DC_BSI,
               cal = DC_CAL, cas = DC_CAS, flnd = DC_FLND, fso = DC_FSO;
  const double S = PAR(PI_S), diff = PAR(PI_DIFF), qco2 = PAR(PI_QCO2);
  const double kcon = DC_SECS_PER_YEAR / 10000;
  const double cnum = rlam * flnd + bsi * (1.0 - flnd);
  const double cden = rlam * flnd - ak * (rlam - bsi);
  const double cfl = flnd * cnum / cden * qco2 / S - bk * (rlam - bsi) / cden;
  const double cfs = (rlam * flnd - ak / (1.0 - flnd) * (rlam - bsi)) * cnum / cden * qco2 / S +
                     rlam * flnd / (1.0 - flnd) * bk * (rlam - bsi) / cden;
  const double kls = bk * rlam * flnd / cden - ak * flnd * cnum / cden * qco2 / S;
  const double keff = kcon * diff;
  const double taubot = (DC_ZBOT * DC_ZBOT) / keff;
  const double taucfs = cas / cfs;
  const double taucfl = cal / cfl;
  const double taudif = (cas * cas) / (csw * csw) * M_PI / keff;
  const double tauksl = (1.0 - flnd) * cas / kls;
  const double taukls = flnd * cal / kls;

  /* lag kernel K(j), j = 1..nrow (E-4); Ker[i] of the reference is K(ns - i) */
  const double tau = taubot / dt;
  double *ker = BS.ker;
  const size_t Mp = HX_BLOCK; /* row stride inside the tile */
  ker[0] = 0.0;
  const double K1v = ker_first(tau);
  ker[1 * Mp] = K1v;
  KerTerm tm1 = ker_term(tau, 1.0), tc = ker_term(tau, 2.0);
  for (int j = 2; j <= C.nrow; ++j) {
    KerTerm tp = ker_term(tau, (double)(j + 1));
    ker[(size_t)j * Mp] = ker_combine(tau, tm1, tc, tp);
    tm1 = tc;
    tc = tp;
  }

  double Cc[4], A[4], B[4];
  Cc[0] = 1.0 / (taucfl * taucfl) + 1.0 / (taukls * taukls) + 2.0 / taucfl / taukls +
          bsi / taukls / tauksl;
  Cc[1] = -1 * bsi / (taukls * taukls) - bsi / taucfl / taukls - bsi / taucfs / taukls -
          (bsi * bsi) / taukls / tauksl;
  Cc[2] = -1 * bsi / (tauksl * tauksl) - 1.0 / taucfs / tauksl - 1.0 / taucfl / tauksl -
          1.0 / taukls / tauksl;
  Cc[3] = 1.0 / (taucfs * taucfs) + (bsi * bsi) / (tauksl * tauksl) + 2.0 * bsi / taucfs / tauksl +
          bsi / taukls / tauksl;
  for (int i = 0; i < 4; i++) Cc[i] = Cc[i] * ((dt * dt) / 12.0);
  const double sqdt = sqrt(dt / taudif);
  B[0] = 1.0 + dt / (2.0 * taucfl) + dt / (2.0 * taukls);
  B[1] = -dt / (2.0 * taukls) * bsi;
  B[2] = -dt / (2.0 * tauksl);
  B[3] = 1.0 + dt / (2.0 * taucfs) + dt / (2.0 * tauksl) * bsi + 2.0 * fso * sqdt;
  A[0] = 1.0 - dt / (2.0 * taucfl) - dt / (2.0 * taukls);
  A[1] = dt / (2.0 * taukls) * bsi;
  A[2] = dt / (2.0 * tauksl);
  A[3] = 1.0 - dt / (2.0 * taucfs) - dt / (2.0 * tauksl) * bsi + K1v * fso * sqdt;
  for (int i = 0; i < 4; i++) {
    B[i] = B[i] + Cc[i];
    A[i] = A[i] + Cc[i];
  }
  /* invert_1d_2x2_matrix, temperature_component.cpp:81-94 */
  const double det = (B[0] * B[3] - B[1] * B[2]);
  const double inv = 1 / det;
  DER(DI_A0) = A[0]; DER(DI_A1) = A[1]; DER(DI_A2) = A[2]; DER(DI_A3) = A[3];
  DER(DI_IB0) = inv * B[3]; DER(DI_IB1) = inv * -1 * B[1]; DER(DI_IB2) = inv * -1 * B[2];
  DER(DI_IB3) = inv * B[0];
  DER(DI_TAUCFL) = taucfl; DER(DI_TAUKLS) = taukls; DER(DI_TAUCFS) = taucfs;
  DER(DI_TAUKSL) = tauksl;
  DER(DI_SQDT_TAUDIF) = sqdt;
  DER(DI_HF_INT) = cas * fso / sqrt(taudif * dt);
  DER(DI_LNQ10) = log(PAR(PI_Q10));
  /* QL = QO = forcing (temperature_component.cpp:462-463), so DelQL = DelQO = dQ and
   * QC1 = dQ (1/cal (1/taucfl + 1/taukls) - bsi/cas/taukls) dt^2/12, QC2 likewise */
  DER(DI_QC1) = (1.0 / cal * (1.0 / taucfl + 1.0 / taukls) - bsi * 1.0 / cas / taukls) * (dt * dt) / 12.0;
  DER(DI_QC2) = (1.0 / cas * (1.0 / taucfs + bsi / tauksl) - 1.0 / cal / tauksl) * (dt * dt) / 12.0;
  DER(DI_INV_UC_CH4) = 1.0 / PAR(PI_UC_CH4);
  DER(DI_INV_TSOIL) = 1.0 / PAR(PI_TSOIL);
  DER(DI_INV_TSTRAT) = 1.0 / PAR(PI_TSTRAT);
  DER(DI_LOG_M0) = hx_log(PAR(PI_M0));
  DER(DI_SQRT_M0) = sqrt(PAR(PI_M0));
}

/* initial pools: ocean_component.cpp:224-260, simpleNbox.cpp:45-81, simpleNbox-runtime.cpp:172 */
__device__ __forceinline__ void setup_state(const HxDev &d, const HxConst &C, const Bases &BS, int m) {
  const double LL_vol_frac = C.vol_LL / (C.vol_LL + C.vol_HL);
  const double HL_vol_frac = 1 - LL_vol_frac;
  const double I_vol_frac = C.vol_IO / (C.vol_IO + C.vol_DO);
  const double D_vol_frac = 1 - I_vol_frac;
  const double pre_s = PAR(PI_PREIND_SURF), pre_id = PAR(PI_PREIND_ID);
  STATE(SI_BOX_LL) = LL_vol_frac * pre_s;
  STATE(SI_BOX_HL) = HL_vol_frac * pre_s;
  STATE(SI_BOX_IO) = I_vol_frac * pre_id;
  STATE(SI_BOX_DO) = D_vol_frac * pre_id;
  STATE(SI_ATMOS) = PAR(PI_C0) * HX_PPMVCO2_TO_PGC;
  STATE(SI_VEG) = PAR(PI_VEG_C0); STATE(SI_DET) = PAR(PI_DET_C0); STATE(SI_SOIL) = PAR(PI_SOIL_C0);
  STATE(SI_PERMAFROST) = PAR(PI_PERMAFROST_C0); STATE(SI_THAWED) = 0.0; STATE(SI_EARTH) = 5500;
  STATE(SI_ALK_HL) = 0.0; STATE(SI_ALK_LL) = 0.0; STATE(SI_H_HL) = 0.0; STATE(SI_H_LL) = 0.0;
  STATE(SI_TEMPFERTS) = 1.0; STATE(SI_F_FROZEN) = 1.0; STATE(SI_CUM_LUC_VA) = 0.0;
  STATE(SI_EOS_VEGC) = PAR(PI_VEG_C0); STATE(SI_MASSTOT) = 0.0; STATE(SI_CUM_PF_CH4) = 0.0;
  STATE(SI_RH_CH4) = 0.0;
  STATE(SI_MAX_TIMESTEP) = HX_OCEAN_MAX_TIMESTEP; STATE(SI_TIMEOUT) = 0.0;
  STATE(SI_LASTFLUX_ANN) = 0.0; STATE(SI_SOLVER_DT) = PAR(PI_DT);
  {
    /* CH4Component::prepareToRun (ch4_component.cpp:137-147): a CH4 constraint at the start
     * date replaces the preindustrial value M0 (OH keeps the parameter: it read M0 earlier) */
    const double c0 = d.scen[(size_t)d.block_scen[blockIdx.x] * C.nrow * SC_STRIDE + SC_C_CH4];
    STATE(SI_CH4) = (c0 == c0) ? c0 : PAR(PI_M0);
  }
  STATE(SI_TLAND) = 0.0; STATE(SI_SST) = 0.0;
  STATE(SI_HEAT_MIXED) = 0.0; STATE(SI_HEAT_INTERIOR) = 0.0; STATE(SI_RF_PREV) = 0.0;
  STATE(SI_BASE_TOT) = 0.0; STATE(SI_BASE_CO2) = 0.0; STATE(SI_BASE_CH4) = 0.0;
  STATE(SI_BASE_N2O) = 0.0;
  STATE(SI_TLAND_WSUM) = 0.0; STATE(SI_TLAND_WCOMP) = 0.0; STATE(SI_DPAST_RAW) = 0.0;
  STATE(SI_TLAND_C) = 0.0; STATE(SI_SST_C) = 0.0;
  if (BS.BF) {
    /* biome-split pools: each biome starts from its own initial pools, the member-level state
     * holds their sums (sum_map order); simpleNbox.cpp:281-301, simpleNbox-runtime.cpp:79, 145 */
    double sv = 0.0, sd = 0.0, ss = 0.0, sp = 0.0;
    for (int k = 0; k < C.n_biomes; ++k) {
      const int ib = C.biome_order[k];
      const double *bp = BS.BP + (size_t)ib * BP_COUNT * HX_BLOCK;
      double *bf = BS.BF + (size_t)ib * BF_COUNT * HX_BLOCK;
      for (int f = 0; f < BF_COUNT; ++f) bf[f * HX_BLOCK] = 0.0;
      sv += (bf[BF_VEG * HX_BLOCK] = __ldg(bp + BP_VEG_C0 * HX_BLOCK));
      sd += (bf[BF_DET * HX_BLOCK] = __ldg(bp + BP_DET_C0 * HX_BLOCK));
      ss += (bf[BF_SOIL * HX_BLOCK] = __ldg(bp + BP_SOIL_C0 * HX_BLOCK));
      sp += (bf[BF_PERMAFROST * HX_BLOCK] = __ldg(bp + BP_PERMAFROST_C0 * HX_BLOCK));
      bf[BF_TEMPFERTS * HX_BLOCK] = 1.0; bf[BF_F_FROZEN * HX_BLOCK] = 1.0;
      bf[BF_X_CO2FERT * HX_BLOCK] = 1.0; bf[BF_X_TFD * HX_BLOCK] = 1.0;
      bf[BF_X_TFS * HX_BLOCK] = 1.0;
    }
    STATE(SI_VEG) = sv; STATE(SI_DET) = sd; STATE(SI_SOIL) = ss; STATE(SI_PERMAFROST) = sp;
    STATE(SI_EOS_VEGC) = sv;
  }
  if (BS.GF) {
    /* N2OComponent / HalocarbonComponent::prepareToRun: N2O starts at N0 (n2o_component.cpp:
     * 141-146), every halocarbon at H0 (halocarbon_component.cpp:160-166); exp(-1 / tau) once */
    BS.GF[GF_N2O * HX_BLOCK] = PAR(PI_N0);
    for (int g = 0; g < HX_NHALO; ++g) {
      const double tau = __ldg(BS.GP + (GP_HALO0 + 5 * g + 0) * HX_BLOCK);
      BS.GF[(GF_HA0 + g) * HX_BLOCK] = __ldg(BS.GP + (GP_HALO0 + 5 * g + 3) * HX_BLOCK);
      BS.GF[(GF_EXPFAC0 + g) * HX_BLOCK] = hx_exp(-(1 / tau));
      BS.GF[(GF_RF0 + g) * HX_BLOCK] = 0.0;
    }
  }
  BS.sst[0] = 0.0;   /* row 0: temp_sst[0] = 0 */
  BS.tland[0] = 0.0;
  d.fail_year[m] = 0;
  d.spinup_steps[m] = 0;
}

/* ======================================================================================== */
/* oceanbox::chem_equilibrate (oceanbox.cpp:382-445): 21-point scan, then
 * boost::math::tools::brent_find_minima(fmin, 2100e-6, 2750e-6, 31) on
 * |flux(alk) - f_target|; alk is left at the LAST point evaluated (oceanbox.cpp:338). */
struct EqBox {
  const HxConst *C;
  ChemK k;
  double carbon, volume, As, CO2, target, alk, h;
  bool ok;
};
__device__ __noinline__ double eq_fmin(EqBox &b, double alk, Work &w) {
  b.alk = alk;
  b.h = 0.0;
  const double pco2 = csys_box(*b.C, b.k, b.carbon, alk, b.volume, b.h, true, b.ok, w);
  return fabs(surface_flux(b.CO2, pco2, 1.0, b.k.G) - b.target);
}
__device__ __noinline__ void chem_equilibrate(EqBox &b, Work &w) {
  const double alk_min = 2100e-6, alk_max = 2750e-6;
  for (double alk1 = alk_min; alk1 <= alk_max; alk1 += (alk_max - alk_min) / 20)
    (void)eq_fmin(b, alk1, w);
  const double tolerance = 2.9802322387695312e-08; /* ldexp(1.0, 1 - 26) = 2^-25 */
  double min = alk_min, max = alk_max;
  double x, wv, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  const double golden = 0.3819660f;
  x = wv = v = max;
  fw = fv = fx = eq_fmin(b, x, w);
  delta2 = delta = 0;
  for (int it = 0; it < 1000; ++it) {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    if (fabs(delta2) > fract1) {
      double r = (x - wv) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double pp = (x - v) * q - (x - wv) * r;
      q = 2 * (q - r);
      if (q > 0) pp = -pp;
      q = fabs(q);
      const double td = delta2;
      delta2 = delta;
      if ((fabs(pp) >= fabs(q * td / 2)) || (pp <= q * (min - x)) || (pp >= q * (max - x))) {
        delta2 = (x >= mid) ? min - x : max - x;
        delta = golden * delta2;
      } else {
        delta = pp / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2))
          delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
      }
    } else {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta)
                                : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    fu = eq_fmin(b, u, w);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = wv; wv = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (wv == x)) {
        v = wv; wv = u;
        fv = fw; fw = fu;
      } else if ((fu <= fv) || (v == x) || (v == wv)) {
        v = u;
        fv = fu;
      }
    }
  }
}

/* Core::run_spinup (core.cpp:394-420) + CarbonCycleSolver::run_spinup
 * (carbon-cycle-solver.cpp:313-370); then the first-year chemistry switch-on of
 * OceanComponent::run (ocean_component.cpp:392-400). */
template <bool BIOMES>
__global__ void __launch_bounds__(HX_BLOCK)
hx_spinup_kernel(const __grid_constant__ HxDev d, const __grid_constant__ HxConst C, int m_offset,
                 int only_member) {
  __shared__ __align__(16) double rk[HX_RK_SLOTS][HX_BLOCK]; /* Runge-Kutta stage derivatives k1..k7 */
  const int m = m_offset + blockIdx.x * HX_BLOCK + threadIdx.x;
  if (m >= d.Mpad) return;
  if (d.status[m] < 0) return;
  if (only_member >= 0 && m != only_member) return;
  const Bases BS = make_bases(d, C, m);
  Member mb;
  load_member(BS, mb);
  const LandPar p = load_landpar(BS);
  Work w = {0, 0, 0, 0, 0, 0};
  ChemRef ck;
  ck.base = nullptr; ck.stride = 0; ck.tid = 0; /* no chemistry during spin-up */
  const double eps_spinup = PAR(PI_EPS_SPINUP);
  int steps = 0;
  if (!(C.flags & HX_FLAG_NO_SPINUP)) {
    mb.S[SI_X_CO2FERT * HX_TILE] = 1.0; mb.S[SI_X_TFD * HX_TILE] = 1.0; mb.S[SI_X_TFS * HX_TILE] = 1.0; mb.S[SI_X_FNEWTHAW * HX_TILE] = 0.0; mb.S[SI_F_FROZEN * HX_TILE] = 1.0;
    mb.luc_e = mb.luc_u = mb.S[SI_X_FFI * HX_TILE] = mb.S[SI_X_DACCS * HX_TILE] = 0.0;
    if (BIOMES)
      for (int ib = 0; ib < C.n_biomes; ++ib) { /* no perturbation in spin-up (:968, 993-997) */
        const Biome b = biome_of(mb, ib);
        b.f(BF_X_CO2FERT) = 1.0; b.f(BF_X_TFD) = 1.0; b.f(BF_X_TFS) = 1.0;
        b.f(BF_X_FNEWTHAW) = 0.0; b.f(BF_F_FROZEN) = 1.0;
      }
    bool spunup = false;
    int step = 0;
    while (!spunup && ++step < C.max_spinup) {
      mb.S[SI_X_FLUXSUM * HX_TILE] = 0.0; mb.timesteps = 0;
      mb.S[SI_X_NPPLUC * HX_TILE] = (mb.S[SI_EOS_VEGC * HX_TILE] - mb.S[SI_CUM_LUC_VA * HX_TILE]) / mb.S[SI_EOS_VEGC * HX_TILE];
      const double o0 = mb.atmos, o1 = mb.veg, o2 = mb.det, o3 = mb.soil, o4 = mb.perm,
                   o5 = mb.thawed, o6 = total_ocean(mb), o7 = mb.earth;
      solver_year<true, false, false, BIOMES>(mb, C, p, ck, &rk[0][threadIdx.x], HX_BLOCK, (double)(step - 1), (double)step, true, w);
      if (mb.status) break;
      double mx = fabs(mb.atmos - o0);
      mx = fmax(mx, fabs(mb.veg - o1)); mx = fmax(mx, fabs(mb.det - o2));
      mx = fmax(mx, fabs(mb.soil - o3)); mx = fmax(mx, fabs(mb.perm - o4));
      mx = fmax(mx, fabs(mb.thawed - o5)); mx = fmax(mx, fabs(total_ocean(mb) - o6));
      mx = fmax(mx, fabs(mb.earth - o7));
      spunup = (mx < eps_spinup);
    }
    steps = step;
    mb.S[SI_RH_CH4 * HX_TILE] = 0.0;           /* record_state in spin-up: simpleNbox.cpp:809-814 */
    mb.S[SI_TEMPFERTS * HX_TILE] = 1.0;
    if (BIOMES)
      for (int ib = 0; ib < C.n_biomes; ++ib) {
        const Biome b = biome_of(mb, ib);
        b.f(BF_RH_CH4) = 0.0; b.f(BF_TEMPFERTS) = 1.0;
      }
  }
  mb.S[SI_EOS_VEGC * HX_TILE] = mb.veg;        /* SimpleNbox::run first call: simpleNbox-runtime.cpp:209-213 */
  if (mb.status == 0) {
    /* chem_equilibrate both surface boxes at SST = 0 and the post-spin-up CO2 */
    const double CO2 = mb.atmos * HX_PGC_TO_PPMVCO2;
    EqBox b;
    b.C = &C; b.ok = true; b.CO2 = CO2;
    b.k = chem_constants(C, 0.0 + HX_MEAN_TOS_TEMP + HX_DT_HL, C.As_HL);
    b.carbon = mb.bHL; b.volume = C.vol_HL; b.As = C.As_HL; b.target = 1.000;
    chem_equilibrate(b, w);
    mb.S[SI_ALK_HL * HX_TILE] = b.alk; mb.S[SI_H_HL * HX_TILE] = b.h;
    b.k = chem_constants(C, 0.0 + HX_MEAN_TOS_TEMP + HX_DT_LL, C.As_LL);
    b.carbon = mb.bLL; b.volume = C.vol_LL; b.As = C.As_LL; b.target = -1.000;
    chem_equilibrate(b, w);
    mb.S[SI_ALK_LL * HX_TILE] = b.alk; mb.S[SI_H_LL * HX_TILE] = b.h;
    if (!b.ok) mb.status = HX_MEMBER_NOROOT;
  }
  store_member(BS, mb);
  d.spinup_steps[m] = steps;
  if (mb.status) {
    d.status[m] = mb.status;
    d.fail_year[m] = C.start_year;
  }
}


/* ======================================================================================== */
/* Slab prepass of the DOECLIM convolution (temperature_component.cpp:488-491, 534-537).
 *
 * Year r needs DPAST2(r) = sum_{i<r} sst[i] K(r-i+1), an O(r) pass over the member's SST history
 * and lag kernel -- streamed from HBM every year that is 16 r bytes per member-year and it
 * dominated both the DRAM traffic and the stall profile of the run kernel.  A work item covers
 * B = HX_SLAB_YEARS consecutive years r_j = base+1+j, and the history rows i < n_pre <= base+1
 * are common to all of them, so ONE pass accumulates the B partial sums
 *     acc[j] = sum_{i<n_pre} sst[i] K(r_j-i+1)
 * with a register window over K that slides one row per history row; each year then only adds
 * its last r - n_pre < B + U terms.  Every acc[j] chain adds the same products in the same
 * (ascending i) order as a year-by-year loop, so results are bit-identical; history traffic
 * drops by B.  (The second convolution of the reference, sum_{i<r} sst[i] K(r-i), is last
 * year's DPAST2 sum plus one term -- see the year loop -- and needs no pass of its own.)
 *
 * U rows are fetched per trip while the B*U FMAs of the previous trip issue. */
template <int B, int U>
__device__ __forceinline__ void conv_prepass(const double *__restrict__ sst,
                                             const double *__restrict__ ker, int base, int n_pre,
                                             double *__restrict__ out) {
  constexpr size_t Hs = HX_BLOCK;
  double acc[B];
#pragma unroll
  for (int j = 0; j < B; ++j) acc[j] = 0.0;
  if (n_pre > 0) {
    /* V[p] = K(base + 2 - i0 - (U-1) + p), p = 0 .. B+U-2: the rows K(r_j - i + 1) that the
     * history rows i = i0 .. i0+U-1 of the current trip meet */
    double V[B + U - 1], s[U], sn[U], vn[U];
    const double *pk = ker + (size_t)(base + 2 - (U - 1)) * Hs;
#pragma unroll
    for (int p = 0; p < B + U - 1; ++p) V[p] = __ldcs(pk + (size_t)p * Hs);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] = __ldcs(sst + (size_t)u * Hs);
      sn[u] = 0.0; vn[u] = 0.0;
    }
    for (int i0 = 0; i0 < n_pre; i0 += U) {
#if HX_CONV_L2_AHEAD
      /* the rows of trip i0 + HX_CONV_L2_AHEAD U on their way from HBM to L2 now: the register
       * prefetch below is one trip deep, enough for L2's latency but not for DRAM's */
      if (i0 + HX_CONV_L2_AHEAD * U < n_pre) {
        const double *fs = sst + (size_t)(i0 + HX_CONV_L2_AHEAD * U) * Hs;
        const double *fk = pk - (size_t)(i0 + HX_CONV_L2_AHEAD * U) * Hs;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(fs + (size_t)u * Hs));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(fk + (size_t)u * Hs));
        }
      }
#endif
      if (i0 + U < n_pre) {
        const double *ns = sst + (size_t)(i0 + U) * Hs;
        const double *nk = pk - (size_t)(i0 + U) * Hs;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          sn[u] = __ldcs(ns + (size_t)u * Hs);
          vn[u] = __ldcs(nk + (size_t)u * Hs);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < B; ++j) acc[j] = HX_CONV_FMA(s[u], V[j + (U - 1) - u], acc[j]);
#pragma unroll
      for (int p = B + U - 2; p >= U; --p) V[p] = V[p - U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        V[u] = vn[u];
        s[u] = sn[u];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < B; ++j) out[(size_t)j * Hs] = acc[j];
}

/* ======================================================================================== */
/* the yearly coupled step, rows r0+1 .. r1 (row = year - start_year).  TRACK = carbon tracking
 * compiled in (a second instantiation: the plain kernel carries none of its code). */
/* failed members report NaN from the failing year on (the reference stops producing output) */
__device__ __noinline__ void nan_fill_rows(const HxDev &d, int nyears, int col, int first, int last) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int s = 0; s < d.n_out; ++s)
    for (int yi = first; yi < last; ++yi) d.out[((size_t)s * nyears + yi) * d.Mpad + col] = nan;
}

/* LAT: the build for ensembles so small that every warp has a scheduler to itself (latency of one
 * warp is all that counts; code size is not): see integrate<.., RKU> */
template <bool TRACK, bool CONSTR, int MINCTAS, bool ALLOUT, bool BIOMES, bool NBP, bool EXACT, bool GAS, bool LAT,
          bool RF4>
__global__ void __launch_bounds__(HX_BLOCK, MINCTAS)
hx_run_kernel(const __grid_constant__ HxDev d, const __grid_constant__ HxConst C, int r0, int r1) {
  /* dynamic shared memory (> 48 KB): scenario slabs | row 0 | chemistry constants | RK stages */
  extern __shared__ __align__(128) unsigned char hx_smem[];
  double (*slab)[(HX_SLAB_YEARS + 1) * SC_STRIDE] =
      reinterpret_cast<double (*)[(HX_SLAB_YEARS + 1) * SC_STRIDE]>(hx_smem);
  double *row0 = reinterpret_cast<double *>(hx_smem + HX_SMEM_ROW0);
  double (*chemk)[HX_BLOCK] = reinterpret_cast<double (*)[HX_BLOCK]>(hx_smem + HX_SMEM_CHEMK);
  double (*rk)[HX_BLOCK] = reinterpret_cast<double (*)[HX_BLOCK]>(hx_smem + HX_SMEM_RK);
  __shared__ __align__(8) uint64_t bars[1];
  const int tid = threadIdx.x;
  const int nyears_total = C.nrow - 1;
  const bool cold = (C.flags & HX_FLAG_COLD_NEWTON) != 0;
  const size_t Mp = d.Mpad;          /* member stride of the output block */
  const size_t Hs = HX_BLOCK;        /* row stride of the tiled history arrays */
  ChemRef ck;
  ck.base = &chemk[0][0]; ck.stride = HX_BLOCK; ck.tid = tid;
  Work w = {0, 0, 0, 0, 0, 0};
  unsigned years_done = 0, failed = 0;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_fence_init();
  }
  __syncthreads();

  /* Persistent CTAs over the work items (tile, slab): the HX_SLAB_YEARS years of one slab for the
   * 128 members of one tile.  A tile's slabs are sequentially dependent (state in S), its slabs
   * may run on different CTAs.  A free CTA claims the LEAST ADVANCED tile that nobody is working
   * on: its first warp scans the tiles' progress words (a warp-wide minimum by shuffles), lane 0
   * takes the tile with a compare-and-swap on its busy word.  All tiles therefore advance
   * together, slab-major -- what the streaming output wants -- and no CTA ever waits for a
   * particular tile: with members sorted so that the members of a tile behave alike, tiles differ
   * in cost by up to 1.5x and a fixed ticket order left the CTA that drew a slow tile's next slab
   * spinning until the previous one was published.  (A static one-CTA-per-tile grid needs
   * ceil(tiles / resident CTAs) full rounds: 2 for 512 tiles on 444 resident CTAs.) */
  const int ntiles = d.Mpad / HX_BLOCK;
  const int nslab = (r1 - r0 + HX_SLAB_YEARS - 1) / HX_SLAB_YEARS;
  /* one word per tile: (slabs of this launch already published << 1) | busy */
  unsigned *state = d.sched + 1;
  unsigned item_no = 0;                        /* items this CTA has processed: mbarrier phase */
  __shared__ int s_tile, s_slab;
  /* equal progress: prefer the tiles after this CTA's own offset, so that CTAs starting together
   * do not all reach for the same tile */
  const unsigned my_off = (unsigned)(((unsigned long long)blockIdx.x * (unsigned)ntiles) / gridDim.x);

  for (;;) {
    if (tid < 32) {
      int got_tile = -1, got_slab = 0;
      for (;;) {
        unsigned best = 0xffffffffu;
        int remaining = 0;
        /* L2-coherent loads (ld.global.cg), four in flight per lane */
        for (int t0 = tid; t0 < ntiles; t0 += 128) {
          unsigned v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = (t0 + 32 * u < ntiles) ? __ldcg(state + t0 + 32 * u) : 0xffffffffu;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const unsigned pr = v[u] >> 1;
            if (v[u] != 0xffffffffu && pr < (unsigned)nslab) {
              remaining = 1;
              const unsigned rot = ((unsigned)(t0 + 32 * u) + (unsigned)ntiles - my_off) % (unsigned)ntiles;
              if (!(v[u] & 1u)) best = min(best, (pr << 20) | rot); /* up to 2^20 tiles, 2047 slabs */
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
          remaining |= __shfl_xor_sync(0xffffffffu, remaining, o);
        }
        if (best == 0xffffffffu) {
          if (!remaining) break;      /* every tile has finished the launch */
          __nanosleep(400);           /* all unfinished tiles are being worked on */
          continue;
        }
        const int t = (int)(((best & 0xfffffu) + my_off) % (unsigned)ntiles);
        const unsigned want = (best >> 20) << 1;
        int ok = 0;
        if (tid == 0 && atomicCAS(state + t, want, want | 1u) == want) {
          /* acquire: the tile's state as its previous slab left it (also drops stale L1 lines) */
          unsigned seen;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(state + t) : "memory");
          ok = 1;
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) { got_tile = t; got_slab = (int)(best >> 20); break; }
      }
      if (tid == 0) { s_tile = got_tile; s_slab = got_slab; }
    }
    __syncthreads();
    const int tile = s_tile;
    if (tile < 0) break;
    const int s = s_slab;
    const int m = tile * HX_BLOCK + tid;
    const double *table = d.scen + (size_t)d.block_scen[tile] * C.nrow * SC_STRIDE;
    /* slab s covers table rows base .. base+HX_SLAB_YEARS (one overlap row for the y-1
     * emissions), serving years base+1 .. base+HX_SLAB_YEARS */
    const int base = r0 + s * HX_SLAB_YEARS;
    const int rend = min(base + HX_SLAB_YEARS, r1);
    if (tid == 0) {
      int rows = HX_SLAB_YEARS + 1;
      if (base + rows > C.nrow) rows = C.nrow - base;
      const uint32_t bytes = (uint32_t)(rows * SC_STRIDE * sizeof(double));
      fence_proxy_async();
      mbar_arrive_expect_tx(&bars[0], bytes);
      bulk_g2s(slab[0], table + (size_t)base * SC_STRIDE, bytes, &bars[0]);
    }
    for (int i = tid; i < SC_STRIDE; i += HX_BLOCK) row0[i] = table[i];
    __syncthreads();
    __threadfence(); /* every thread orders its state loads after the acquire above */

    const bool lane_ok = (d.status[m] == 0);
    /* the output block is in API member order (what callers, copies and exchanges see); the
     * engine orders members internally (by scenario, then so that the members of a warp behave
     * alike), so a member's column is looked up once per work item.  -1: padding lane. */
    const int mo = d.api_of_dev[m];
    Bases BS = make_bases(d, C, m);
    double *const gS = BS.S;
    double *const gBF = BS.BF;
    /* the block of shared memory behind the kernel's own arrays holds the tile's state S -- or,
     * in the biome builds, the per-biome pools and factors BF (up to three biomes fit): the
     * biome loops of every sub-step and stash read and write them far more often than S */
    constexpr bool S_IN_SMEM = smem_state<MINCTAS>() && !BIOMES;
    const bool bf_in_smem = smem_state<MINCTAS>() && BIOMES && C.n_biomes * BF_COUNT <= HX_HOT_COUNT + SI_COUNT - SI_REG_COUNT;
    if (S_IN_SMEM) {
      /* [the hot stretch of P | D][the state without its register-resident fields]: the block
       * is as large as the whole state was, the fourteen fields that only load_member /
       * store_member touch made room for the constants every sub-step and stash reads */
      double *smH = reinterpret_cast<double *>(hx_smem + HX_SMEM_RUN_BYTES) + tid;
      double *smS = smH + ((LAT ? PD_COUNT : HX_HOT_COUNT) - SI_REG_COUNT) * HX_BLOCK; /* field i at smS[i * HX_BLOCK], i >= SI_REG_COUNT */
      if (LAT) { /* every parameter and derived constant: the year loop reads none of them from L2 */
        const double *gP = BS.P;
#pragma unroll 8
        for (int i = 0; i < PD_COUNT; ++i) smH[i * HX_BLOCK] = __ldg(gP + i * HX_BLOCK);
        BS.P = smH;
        BS.D = smH + PI_COUNT * HX_BLOCK;
        BS.H = smH + HX_HOT_FIRST * HX_BLOCK;
        BS.psm = true;
      } else {
        const double *gH = BS.H;
#pragma unroll
        for (int i = 0; i < HX_HOT_COUNT; ++i) smH[i * HX_BLOCK] = __ldg(gH + i * HX_BLOCK);
        BS.H = smH;
      }
#pragma unroll 8
      for (int i = SI_REG_COUNT; i < SI_COUNT; ++i) smS[i * HX_BLOCK] = gS[i * HX_BLOCK];
      BS.S = smS;
    }
    if (bf_in_smem) {
      double *smB = reinterpret_cast<double *>(hx_smem + HX_SMEM_RUN_BYTES) + tid;
      const int nf = C.n_biomes * BF_COUNT;
#pragma unroll 4
      for (int i = 0; i < nf; ++i) smB[i * HX_BLOCK] = gBF[i * HX_BLOCK];
      BS.BF = smB;
    }
    /* history rows [0, n_pre) of the DOECLIM convolution, for all years of the slab at once */
    const int n_pre = ((base + 1) / HX_CONV_UNROLL) * HX_CONV_UNROLL;
    conv_prepass<HX_SLAB_YEARS, HX_CONV_UNROLL>(BS.sst, BS.ker, base, n_pre, BS.conv);
    const LandPar p = load_landpar(BS);
    Member mb;
    if (lane_ok) load_member(BS, mb);
    else mb.status = -1;
    /* scratch rows of the per-stash outputs: only the all-output builds know them */
    mb.X = (ALLOUT && d.X) ? d.X + (size_t)tile * XS_COUNT * HX_BLOCK + tid : nullptr;
    mbar_wait(&bars[0], item_no & 1u);
    ++item_no;
    const double *sl = slab[0];
    /* Year barrier.  The year body is ~110 KB of SASS, far more than the instruction caches
     * hold; left alone the CTA's four warps drift apart and each streams the loop from L2 on
     * its own (ncu: 2.1 of 8.7 stall cycles per issue are instruction fetch).  One CTA barrier
     * per simulated year keeps them in the same stretch of code so that they share the fetches
     * (36.0 vs 39.1 ms).  Every thread arrives rend - base times per item: members still
     * running at the top of each year, members that stopped early (failed, or padding lanes)
     * in the drain loop below.  The arrivals come from divergent code, hence the non-aligned
     * barrier.sync on a named barrier of its own. */
    unsigned char ycnt[HX_SLAB_YEARS]; /* stashes recorded up to the end of each year (TRACK) */
    if (TRACK) {
#pragma unroll
      for (int j = 0; j < HX_SLAB_YEARS; ++j) ycnt[j] = 0;
      mb.REC = d.REC + (size_t)s * d.rec_slab_stride + 2 * ((size_t)tile * HX_BLOCK + tid);
      mb.rec_stride = (size_t)d.Mpad;
      mb.rec_n = 0;
    }
    int r = base + 1;
    int bars_done = 0; /* barriers this thread has arrived at in this work item */
    const bool entered = (mb.status == 0);
#define HX_OH_LOADS(T)                                                                              \
  T oh_m0 = PAR_AHEAD(BS.P + PI_M0 * HX_BLOCK), oh_cch4 = PAR_AHEAD(BS.P + PI_CCH4 * HX_BLOCK),          \
    oh_cnox = PAR_AHEAD(BS.P + PI_CNOX * HX_BLOCK), oh_cco = PAR_AHEAD(BS.P + PI_CCO * HX_BLOCK),        \
    oh_cnmvoc = PAR_AHEAD(BS.P + PI_CNMVOC * HX_BLOCK), oh_toh0 = PAR_AHEAD(BS.P + PI_TOH0 * HX_BLOCK),  \
    oh_logm0 = PAR_AHEAD(BS.D + DI_LOG_M0 * HX_BLOCK), oh_iuc = PAR_AHEAD(BS.D + DI_INV_UC_CH4 * HX_BLOCK), \
    oh_itsoil = PAR_AHEAD(BS.D + DI_INV_TSOIL * HX_BLOCK), oh_itstrat = PAR_AHEAD(BS.D + DI_INV_TSTRAT * HX_BLOCK);
#if HX_OH_AHEAD == 2
    HX_OH_LOADS(double)
#endif
    if (entered) {
      for (; r <= rend; ++r) {
#if HX_OH_AHEAD == 1
        /* the ten constants of the OH / CH4 block are the year's first reads from L2, and nothing
         * else can run until they arrive: requested BEFORE the year barrier (volatile loads, so
         * that they stay there), their latency passes while the warp waits for the others */
        HX_OH_LOADS(const double)
#elif HX_OH_AHEAD == 2
        /* requested a phase earlier still: before the loop for its first year, after the solver
         * for the next one (below) */
#else
#define oh_m0 PAR(PI_M0)
#define oh_cch4 PAR(PI_CCH4)
#define oh_cnox PAR(PI_CNOX)
#define oh_cco PAR(PI_CCO)
#define oh_cnmvoc PAR(PI_CNMVOC)
#define oh_toh0 PAR(PI_TOH0)
#define oh_logm0 DER(DI_LOG_M0)
#define oh_iuc DER(DI_INV_UC_CH4)
#define oh_itsoil DER(DI_INV_TSOIL)
#define oh_itstrat DER(DI_INV_TSTRAT)
#endif
#if HX_YEAR_SYNC
        if ((r - base - 1) % HX_YEAR_SYNC_EVERY == 0) { year_barrier(); ++bars_done; }
#endif
        {
        const double *sc = sl + (size_t)(r - base) * SC_STRIDE;     /* year y */
        const double *scm1 = sl + (size_t)(r - 1 - base) * SC_STRIDE; /* year y-1 */
        const int y = C.start_year + r;

        /* --- OH, CH4: oh_component.cpp:137-174, ch4_component.cpp:152-199 --- */
        {
          const double M0 = oh_m0;
          const double previous_ch4 = STATE(SI_CH4);
          double toh = 0.0;
          if (previous_ch4 != M0) {
            /* log(previous_ch4) was taken last year for the ozone burden (below) and log(M0) by
             * the set-up kernel: same function, same argument, same bits */
            const double lprev = (r == 1) ? hx_log(previous_ch4) : STATE(SI_LOG_CH4);
            const double a = oh_cch4 * ((1.0 * lprev) - oh_logm0);
            const double b = oh_cnox * ((1.0 * sc[SC_NOX]) - row0[SC_NOX]);
            const double c = oh_cco * ((1.0 * sc[SC_CO]) - row0[SC_CO]);
            const double dd = oh_cnmvoc * ((1.0 * sc[SC_NMVOC]) - row0[SC_NMVOC]);
            toh = a + b + c + dd;
          }
          const double tau_oh = oh_toh0 * hx_exp(-toh);
          const double rh_ch4_tg = mb.S[SI_RH_CH4 * HX_TILE] * (1000.0 * 16.04 / 12.01);
          const double emisTocon = (sc[SC_CH4_E] + rh_ch4_tg + sc[SC_CH4N]) * oh_iuc;
          const double soil_sink = previous_ch4 * oh_itsoil;
          const double strat_sink = previous_ch4 * oh_itstrat;
          const double oh_sink = previous_ch4 / tau_oh;
          const double dCH4 = emisTocon - soil_sink - strat_sink - oh_sink;
          double ch4_new = previous_ch4 + dCH4;
          if (CONSTR) { /* concentration-forced year: ch4_component.cpp:156-158 */
            const double c = sc[SC_C_CH4];
            if (c == c) ch4_new = c;
          }
          STATE(SI_CH4) = ch4_new;
        }

        /* --- N2OComponent::run (n2o_component.cpp:150-191) and the 26 HalocarbonComponent::run
         * (halocarbon_component.cpp:181-229) with this member's own parameters --- */
        double n2o_conc = sc[SC_N2O];
        if (GAS) {
          const double *gt = d.scen_gas + ((size_t)d.block_scen[tile] * C.nrow + r) * HX_GAS_COLS;
          const double N0 = PAR(PI_N0);
          const double previous_n2o = BS.GF[GF_N2O * HX_BLOCK];
          const double tau_n2o = __ldg(BS.GP + GP_TN2O0 * HX_BLOCK) * hx_pow(previous_n2o / N0, -0.05);
          const double current_n2oem = __ldg(gt + 0) + __ldg(gt + 1);
          const double dN2O = current_n2oem / __ldg(BS.GP + GP_UC_N2O * HX_BLOCK) - previous_n2o / tau_n2o;
          n2o_conc = previous_n2o + dN2O;
          BS.GF[GF_N2O * HX_BLOCK] = n2o_conc;
#pragma unroll 2
          for (int g = 0; g < HX_NHALO; ++g) {
            const double *gp = BS.GP + (GP_HALO0 + 5 * g) * HX_BLOCK;
            const double tau = __ldg(gp), rho = __ldg(gp + HX_BLOCK), delta = __ldg(gp + 2 * HX_BLOCK),
                         mm = __ldg(gp + 4 * HX_BLOCK);
            const double expfac = BS.GF[(GF_EXPFAC0 + g) * HX_BLOCK];
            const double emissMol = __ldg(gt + 2 + g) / mm * 1.0;
            const double concDeltaEmiss = emissMol / (0.1 * 1.8);
            const double Ha = BS.GF[(GF_HA0 + g) * HX_BLOCK] * expfac + concDeltaEmiss * tau * (1.0 - expfac);
            BS.GF[(GF_HA0 + g) * HX_BLOCK] = Ha;
            const double rf_unadjusted = rho * Ha;
            BS.GF[(GF_RF0 + g) * HX_BLOCK] = rf_unadjusted + delta * rf_unadjusted;
          }
        }

        SlowPar sp = {0.0, 1.0, 0.0, 0.0, 0.0};
        double th_in = 0.0, th_out = 0.0; /* land-temperature rows entering / leaving the window */

        /* --- OceanComponent::run: ocean_component.cpp:356-407 --- */
        mb.S[SI_X_FLUXSUM * HX_TILE] = 0.0;
        if (ALLOUT && mb.X) { mb.X[XS_UPTAKE_HL * HX_TILE] = 0.0; mb.X[XS_UPTAKE_LL * HX_TILE] = 0.0; }
        mb.timesteps = 0;
        {
          /* getData(sst): DOECLIM's, or the lo_warming_ratio one (temperature_component.cpp:
           * 612-622) */
          const double sst = CONSTR ? STATE(SI_SST_C) : STATE(SI_SST);
          const ChemG g = chem_constants2(C, sst, ck.base + ck.tid, ck.stride);
          mb.gHL = g.gHL; mb.gLL = g.gLL;
          const Csys2Out o = csys_solve2(ck.base + ck.tid, ck.stride, C.bor, mb.bHL, mb.bLL,
                                         STATE(SI_ALK_HL), STATE(SI_ALK_LL), C.inv_vol_HL, C.inv_vol_LL,
                                         STATE(SI_H_HL), STATE(SI_H_LL), cold);
          mb.pco2HL = o.pco2[0]; mb.pco2LL = o.pco2[1];
          STATE(SI_H_HL) = o.h[0]; STATE(SI_H_LL) = o.h[1];
          w.newton_it += o.iters; w.newton_calls += 2;
          if (!o.ok) mb.status = HX_MEMBER_NOROOT;
        }

        /* --- SimpleNbox::run + slowparameval: simpleNbox-runtime.cpp:206-227, 945-1072 --- */
        {
          const double tland = CONSTR ? STATE(SI_TLAND_C) : STATE(SI_TLAND); /* getData(land_tas) */
          /* (requesting these seven values ahead of the chemistry was tried: +0.8 % -- they stay
           * live across its two calls) */
          if (!BIOMES) {
            sp.beta = LP_BETA(p); sp.wf = LP_WF(p); sp.lnq10 = LP_LNQ10(p); sp.pf_mu = LP_PF_MU(p);
            sp.pf_sigma = LP_PF_SIGMA(p);
          }
          if (r - 2 >= 1) th_in = BS.tland[(size_t)(r - 2) * Hs];
          if (r - 202 >= 1) th_out = BS.tland[(size_t)(r - 202) * Hs];
          const double wf = sp.wf; /* 1 with biomes: they weight the window mean themselves */
          BS.tland[(size_t)r * Hs] = tland; /* Tland_record[y] = land tas of year y-1 */
          if (TRACK) {
            /* tracking starts in year tracking_date (simpleNbox-runtime.cpp:215-220,
             * ocean_component.cpp:358-366); the ocean gets this year's copy of the atmosphere's
             * source map (set_atmosphere_sources, :225) */
            mb.trk = (y >= C.tracking_date);
          }
          mb.luc_e = scm1[SC_LUC_E]; mb.luc_u = scm1[SC_LUC_U];
          mb.S[SI_X_FFI * HX_TILE] = scm1[SC_FFI]; mb.S[SI_X_DACCS * HX_TILE] = scm1[SC_DACCS];
          if (CONSTR) {
            mb.S[SI_X_C_CO2 * HX_TILE] = sc[SC_C_CO2];
            mb.S[SI_X_C_NBP0 * HX_TILE] = scm1[SC_C_NBP];
            mb.S[SI_X_C_NBP1 * HX_TILE] = sc[SC_C_NBP];
          }
          mb.neg |= (mb.luc_e < 0.0) | (mb.luc_u < 0.0) | (mb.S[SI_X_FFI * HX_TILE] < 0.0) | (mb.S[SI_X_DACCS * HX_TILE] < 0.0);
          /* Tland_rm: for (i = t-200; i < t; i++) Tland_rm += Tland_record.get(i) * wf; /= 200
           * (:1041-1050).  Keys below the first record (start+1) extrapolate flat to it and
           * that record is exactly 0, so the window is sum_{k = max(1, r-201)}^{r-2} hist[k] wf.
           * It is carried as a compensated (Neumaier) running sum: one term enters and one
           * leaves per year instead of re-reading 200 history rows. */
          double window = 0.0;
          if (r >= 2) {
            double wsum = STATE(SI_TLAND_WSUM), wcomp = STATE(SI_TLAND_WCOMP);
            double add = 0.0, sub = 0.0;
            if (r - 2 >= 1) add = th_in * wf;
            if (r - 202 >= 1) sub = -(th_out * wf);
            double t1 = wsum + add;
            wcomp += (fabs(wsum) >= fabs(add)) ? ((wsum - t1) + add) : ((add - t1) + wsum);
            wsum = t1;
            t1 = wsum + sub;
            wcomp += (fabs(wsum) >= fabs(sub)) ? ((wsum - t1) + sub) : ((sub - t1) + wsum);
            wsum = t1;
            window = (wsum + wcomp) / 200;
            STATE(SI_TLAND_WSUM) = wsum; STATE(SI_TLAND_WCOMP) = wcomp;
          }
          /* log(CO2 / C0) of the atmosphere at the start of the year: taken at the end of last
           * year (below), where the forcing needs the same number */
          const double lco2 = (r == 1) ? hx_log((mb.atmos * HX_PGC_TO_PPMVCO2) / LP_C0(p)) : STATE(SI_LOG_CO2R);
          if (BIOMES) slow_params_biomes(mb, C, p, tland, r == 1, window, lco2);
          else slow_params(mb, sp, tland, r == 1, window, lco2);
        }

#if HX_YEAR_SYNC >= 2
        year_barrier(); ++bars_done; /* the warps enter the solver together */
#endif
        /* --- CarbonCycleSolver::run --- */
        solver_year<false, TRACK, CONSTR, BIOMES, NBP, NBP || EXACT, LAT>(mb, C, p, ck, &rk[0][tid], HX_BLOCK, (double)(y - 1), (double)y, cold, w);
        if (mb.status) {
          d.status[m] = mb.status;
          d.fail_year[m] = y;
          break;
        }
#if HX_YEAR_SYNC >= 3
        year_barrier(); ++bars_done; /* ... and leave it together */
#endif
#if HX_FORC_AHEAD
        /* the forcing's and DOECLIM's member constants, requested as soon as the solver is done:
         * record_state, the logarithm pair and the forcing sum hide their latency */
        const double fa_c0 = PAR_AHEAD(BS.P + PI_C0 * HX_BLOCK), fa_m0 = PAR_AHEAD(BS.P + PI_M0 * HX_BLOCK),
                     fa_aero = PAR_AHEAD(BS.P + PI_AERO * HX_BLOCK), fa_vol = PAR_AHEAD(BS.P + PI_VOL * HX_BLOCK),
                     fa_dco2 = PAR_AHEAD(BS.P + PI_DELTA_CO2 * HX_BLOCK), fa_dch4 = PAR_AHEAD(BS.P + PI_DELTA_CH4 * HX_BLOCK),
                     fa_dn2o = PAR_AHEAD(BS.P + PI_DELTA_N2O * HX_BLOCK), fa_rbc = PAR_AHEAD(BS.P + PI_RHO_BC * HX_BLOCK),
                     fa_roc = PAR_AHEAD(BS.P + PI_RHO_OC * HX_BLOCK), fa_rso2 = PAR_AHEAD(BS.P + PI_RHO_SO2 * HX_BLOCK),
                     fa_rnh3 = PAR_AHEAD(BS.P + PI_RHO_NH3 * HX_BLOCK), fa_sqm0 = PAR_AHEAD(BS.D + DI_SQRT_M0 * HX_BLOCK);
        const double dc_qc1 = PAR_AHEAD(BS.D + DI_QC1 * HX_BLOCK), dc_qc2 = PAR_AHEAD(BS.D + DI_QC2 * HX_BLOCK),
                     dc_a0 = PAR_AHEAD(BS.D + DI_A0 * HX_BLOCK), dc_a1 = PAR_AHEAD(BS.D + DI_A1 * HX_BLOCK),
                     dc_a2 = PAR_AHEAD(BS.D + DI_A2 * HX_BLOCK), dc_a3 = PAR_AHEAD(BS.D + DI_A3 * HX_BLOCK),
                     dc_ib0 = PAR_AHEAD(BS.D + DI_IB0 * HX_BLOCK), dc_ib1 = PAR_AHEAD(BS.D + DI_IB1 * HX_BLOCK),
                     dc_ib2 = PAR_AHEAD(BS.D + DI_IB2 * HX_BLOCK), dc_ib3 = PAR_AHEAD(BS.D + DI_IB3 * HX_BLOCK),
                     dc_sqdt = PAR_AHEAD(BS.D + DI_SQDT_TAUDIF * HX_BLOCK), dc_hfint = PAR_AHEAD(BS.D + DI_HF_INT * HX_BLOCK),
                     dc_k1 = ldg_pinned(BS.ker + Hs); /* K(1): written by the set-up kernel, read-only here */
#else
#define fa_c0 LP_C0(p)
#define fa_m0 PAR(PI_M0)
#define fa_aero PAR(PI_AERO)
#define fa_vol PAR(PI_VOL)
#define fa_dco2 PAR(PI_DELTA_CO2)
#define fa_dch4 PAR(PI_DELTA_CH4)
#define fa_dn2o PAR(PI_DELTA_N2O)
#define fa_rbc PAR(PI_RHO_BC)
#define fa_roc PAR(PI_RHO_OC)
#define fa_rso2 PAR(PI_RHO_SO2)
#define fa_rnh3 PAR(PI_RHO_NH3)
#define fa_sqm0 DER(DI_SQRT_M0)
#define dc_qc1 DER(DI_QC1)
#define dc_qc2 DER(DI_QC2)
#define dc_a0 DER(DI_A0)
#define dc_a1 DER(DI_A1)
#define dc_a2 DER(DI_A2)
#define dc_a3 DER(DI_A3)
#define dc_ib0 DER(DI_IB0)
#define dc_ib1 DER(DI_IB1)
#define dc_ib2 DER(DI_IB2)
#define dc_ib3 DER(DI_IB3)
#define dc_sqdt DER(DI_SQDT_TAUDIF)
#define dc_hfint DER(DI_HF_INT)
#define dc_k1 BS.ker[Hs]
#endif
#if HX_OH_AHEAD == 2
        HX_OH_LOADS() /* next year's */
#endif
        /* record_state: simpleNbox.cpp:789-840 */
        if (BIOMES) {
          for (int ib = 0; ib < C.n_biomes; ++ib) {
            const Biome b = biome_of(mb, ib);
            double npp, rh_fda, rh_fsa, rh_co2, rh_ch4v;
            biome_fluxes(mb, b, npp, rh_fda, rh_fsa, rh_co2, rh_ch4v);
            b.f(BF_RH_CH4) = rh_ch4v;
            b.f(BF_TEMPFERTS) = b.f(BF_X_TFS);
          }
          mb.S[SI_RH_CH4 * HX_TILE] = biome_sum(mb, C, BF_RH_CH4);
        } else {
          double npp, rh_fda, rh_fsa, rh_co2, rh_ch4v;
          land_fluxes<false>(mb, p, npp, rh_fda, rh_fsa, rh_co2, rh_ch4v);
          mb.S[SI_RH_CH4 * HX_TILE] = rh_ch4v;
          mb.S[SI_TEMPFERTS * HX_TILE] = mb.S[SI_X_TFS * HX_TILE];
        }
        const double CO2_conc = mb.atmos * HX_PGC_TO_PPMVCO2;
        mb.neg |= (CO2_conc < 0.0);
        if (mb.neg) {
          mb.status = HX_MEMBER_NEGATIVE;
          d.status[m] = mb.status;
          d.fail_year[m] = y;
          break;
        }

        /* --- OzoneComponent::run (o3_component.cpp:126-146) + ForcingComponent::run --- */
        const double ch4 = STATE(SI_CH4);
        /* the year's two remaining logarithms as one interleaved pair; both are needed again at
         * the top of next year (OH lifetime, CO2 fertilisation) and travel in the state */
        const HxPair lg = hx_log_x2(ch4, CO2_conc / fa_c0);
        STATE(SI_LOG_CH4) = lg.a;
        STATE(SI_LOG_CO2R) = lg.b;
        const double o3 = (5 * lg.a) + (0.125 * sc[SC_NOX]) + (0.0011 * sc[SC_CO]) +
                          (0.0033 * sc[SC_NMVOC]);
        double rf_tot = 0.0, rf_co2 = 0.0, rf_ch4 = 0.0, rf_n2o = 0.0;
        if (y >= C.baseyear) {
          ForcPar fp;
          /* preindustrial CH4 / N2O as the CH4 and N2O components hold them after prepareToRun:
           * a start-date constraint replaces the parameter (ch4_component.cpp:141-146,
           * n2o_component.cpp:141-146; row 0 of the N2O series is N0 by construction) */
          fp.C0 = fa_c0; fp.M0 = fa_m0; fp.N0 = GAS ? PAR(PI_N0) : row0[SC_N2O]; fp.aero = fa_aero;
          /* square roots of member or scenario constants come from the set-up kernel / the
           * scenario table (sqrt is correctly rounded everywhere: the same doubles) */
          fp.sqM0 = fa_sqm0;
          fp.sqN0 = GAS ? sqrt(fp.N0) : row0[SC_SQRT_N2O];
          fp.sqNa = GAS ? sqrt(n2o_conc) : sc[SC_SQRT_N2O];
          fp.ln_co2 = lg.b;
          if (CONSTR) {
            const double c0 = row0[SC_C_CH4];
            if (c0 == c0) { fp.M0 = c0; fp.sqM0 = sqrt(c0); }
          }
          fp.vol = fa_vol; fp.delta_co2 = fa_dco2; fp.delta_ch4 = fa_dch4;
          fp.delta_n2o = fa_dn2o; fp.rho_bc = fa_rbc; fp.rho_oc = fa_roc;
          fp.rho_so2 = fa_rso2; fp.rho_nh3 = fa_rnh3;
          double fco2, fch4, fn2o;
          double F = GAS ? forcing_total<HX_BLOCK>(fp, sc, n2o_conc, BS.GF + GF_RF0 * HX_BLOCK, CO2_conc, ch4,
                                                   o3, fco2, fch4, fn2o, mb.status)
                         : forcing_total<1>(fp, sc, n2o_conc, sc + SC_HALO0, CO2_conc, ch4, o3, fco2,
                                            fch4, fn2o, mb.status);
          if (CONSTR) { /* user-supplied total forcing: forcing_component.cpp:498-505 */
            const double c = sc[SC_C_RFTOT];
            if (c == c) F = c;
          }
          if (y == C.baseyear) {
            STATE(SI_BASE_TOT) = F; STATE(SI_BASE_CO2) = fco2; STATE(SI_BASE_CH4) = fch4;
            STATE(SI_BASE_N2O) = fn2o;
          }
          rf_tot = F - STATE(SI_BASE_TOT); rf_co2 = fco2 - STATE(SI_BASE_CO2);
          rf_ch4 = fch4 - STATE(SI_BASE_CH4); rf_n2o = fn2o - STATE(SI_BASE_N2O);
          if (mb.status) {
            d.status[m] = mb.status;
            d.fail_year[m] = y;
            break;
          }
        }

        /* --- TemperatureComponent::run: temperature_component.cpp:417-557 (tstep = r > 0) --- */
        double tas, heatflux, tland_new, sst_new, hf_mixed_out, hf_int_out, ocean_tas_out, gmst_out;
        {
          const double dt = 1.0, bsi = DC_BSI, cal = DC_CAL, cas = DC_CAS, flnd = DC_FLND,
                       fso = DC_FSO;
          const double tland = STATE(SI_TLAND), sst = STATE(SI_SST), rf_prev = STATE(SI_RF_PREV);
          const double dQ = rf_tot - rf_prev;
          const double QC1 = dQ * dc_qc1;
          const double QC2 = dQ * dc_qc2;
          double DQ1 = 0.5 * dt / cal * (rf_tot + rf_prev);
          double DQ2 = 0.5 * dt / cas * (rf_tot + rf_prev);
          DQ1 = DQ1 + QC1;
          DQ2 = DQ2 + QC2;
          /* the two history convolutions (E-4):
           *   DPAST2   = sum_{i<=t} sst[i] K(t-i+1)   (:488-491; the i = t term is 0)
           *   interior = sum_{i<t}  sst[i] K(t-i)     (:534-537)
           * interior(t) adds, oldest first, exactly the products of DPAST2(t-1) and then the
           * one term sst[t-1] K(1): it is last year's unscaled sum plus one FMA.  DPAST2 takes
           * the rows i < n_pre from the slab prepass and adds its last rows here, oldest first
           * like the reference. */
          const double hint = HX_CONV_FMA(sst, dc_k1, STATE(SI_DPAST_RAW));
          double DPAST2 = BS.conv[(size_t)(r - base - 1) * Hs];
          {
            const double *ps = BS.sst + (size_t)n_pre * Hs;          /* sst[i], i ascending */
            const double *pk = BS.ker + (size_t)(r - n_pre + 1) * Hs; /* K(r - i + 1), descending */
#pragma unroll 4
            for (int i = n_pre; i < r; ++i) {
              DPAST2 = HX_CONV_FMA(*ps, *pk, DPAST2);
              ps += Hs;
              pk -= Hs;
            }
          }
          STATE(SI_DPAST_RAW) = DPAST2;
          DPAST2 = DPAST2 * fso * dc_sqdt;
          const double DPAST1 = 0.0;
          const double DTEAUX1 = dc_a0 * tland + dc_a1 * sst;
          const double DTEAUX2 = dc_a2 * tland + dc_a3 * sst;
          double TL = dc_ib0 * (DQ1 + DPAST1 + DTEAUX1) +
                      dc_ib1 * (DQ2 + DPAST2 + DTEAUX2);
          double TS = dc_ib2 * (DQ1 + DPAST1 + DTEAUX1) +
                      dc_ib3 * (DQ2 + DPAST2 + DTEAUX2);
          tas = flnd * TL + (1.0 - flnd) * bsi * TS;
          if (CONSTR) {
            /* user-supplied global temperature (:510-525): overwrite, then back-calculate the
             * land and sea-surface values that go into the histories */
            const double c = sc[SC_C_TAS];
            if (c == c) {
              tas = c;
              TL = (tas - (1.0 - flnd) * bsi * TS) / flnd;
              TS = (tas - flnd * TL) / ((1.0 - flnd) * bsi);
            }
          }
          const double hf_mixed = cas * (TS - sst);
          const double hf_int = dc_hfint * (2.0 * TS - hint);
          STATE(SI_HEAT_MIXED) = STATE(SI_HEAT_MIXED) + hf_mixed * (C.powtoheat * dt);
          STATE(SI_HEAT_INTERIOR) = STATE(SI_HEAT_INTERIOR) + hf_int * (fso * C.powtoheat * dt);
          heatflux = hf_mixed + fso * hf_int;
          hf_mixed_out = hf_mixed; hf_int_out = hf_int;
          tland_new = TL;
          sst_new = TS;
          STATE(SI_TLAND) = TL;
          STATE(SI_SST) = TS;
          STATE(SI_RF_PREV) = rf_tot;
          BS.sst[(size_t)r * Hs] = TS;
          ocean_tas_out = bsi * TS;
          gmst_out = DC_FLND * TL + (1.0 - DC_FLND) * TS; /* always DOECLIM's own (:741-742) */
          if (CONSTR) {
            /* user-provided land-ocean warming ratio (:722-739): the land, ocean-air and sea-
             * surface temperatures other components and callers see are re-derived from the
             * global mean; DOECLIM's own state (above) is untouched */
            const double lo = PAR(PI_LO_RATIO);
            if (lo != 0.0) {
              const double oa = tas / ((lo * flnd) + (1 - flnd));
              tland_new = oa * lo;
              sst_new = oa / bsi;
              ocean_tas_out = oa;
            }
            STATE(SI_TLAND_C) = tland_new;
            STATE(SI_SST_C) = sst_new;
          }
        }
        ++years_done;

        if (TRACK) ycnt[r - base - 1] = (unsigned char)mb.rec_n;

        /* --- outputs (record_state / getData of each component) --- */
        const int yi = r - 1;
#define EMIT(id, val)                                                              \
  do {                                                                             \
    const int slot_ = d.out_slot[id];                                              \
    if (slot_ >= 0) d.out[((size_t)slot_ * nyears_total + yi) * Mp + mo] = (val);  \
  } while (0)
        EMIT(OUT_CO2, CO2_conc);
        EMIT(OUT_TAS, tas);
        if (ALLOUT || RF4) { /* R's default fetchvars are CO2, RF_tot, RF_CO2, Tgav (R/messages.R:47-52):
                                the RF4 builds are the CO2 / Tgav-only ones plus these two rows
                                (26.4 against 28.5 ms in the all-output build; the CO2 / Tgav-only
                                build itself would lose 0.5 % carrying them) */
        EMIT(OUT_RF_TOT, rf_tot);
        EMIT(OUT_RF_CO2, rf_co2);
        }
        if (ALLOUT) { /* the default outputs have a build without the other thirty:
                         less code in the year body is worth 5 % (36.5 -> 34.6 ms) */
        EMIT(OUT_HEATFLUX, heatflux);
        EMIT(OUT_OCEAN_C, mb.bDO + mb.bIO + mb.bLL + mb.bHL);
        if (d.out_slot[OUT_HL_PH] >= 0) EMIT(OUT_HL_PH, -hx_log10(mb.S[SI_H_HL * HX_TILE]));
        EMIT(OUT_ATMOS_C, mb.atmos);
        EMIT(OUT_SST, sst_new);
        EMIT(OUT_PERMAFROST_C, mb.perm);
        EMIT(OUT_CH4, ch4);
        EMIT(OUT_N2O, n2o_conc);
        EMIT(OUT_O3, o3);
        EMIT(OUT_LAND_TAS, tland_new);
        EMIT(OUT_VEG_C, mb.veg);
        EMIT(OUT_DETRITUS_C, mb.det);
        EMIT(OUT_SOIL_C, mb.soil);
        EMIT(OUT_THAWEDP_C, mb.thawed);
        EMIT(OUT_EARTH_C, mb.earth);
        EMIT(OUT_NBP, mb.S[SI_X_NBP * HX_TILE]);
        EMIT(OUT_OCEAN_UPTAKE, mb.S[SI_X_FLUXSUM * HX_TILE]);
        if (d.out_slot[OUT_LL_PH] >= 0) EMIT(OUT_LL_PH, -hx_log10(mb.S[SI_H_LL * HX_TILE]));
        EMIT(OUT_PCO2_HL, mb.pco2HL);
        EMIT(OUT_PCO2_LL, mb.pco2LL);
        EMIT(OUT_CARBON_HL, mb.bHL);
        EMIT(OUT_CARBON_LL, mb.bLL);
        EMIT(OUT_CARBON_IO, mb.bIO);
        EMIT(OUT_CARBON_DO, mb.bDO);
        EMIT(OUT_RF_CH4, rf_ch4);
        EMIT(OUT_RF_N2O, rf_n2o);
        EMIT(OUT_RH_CH4, mb.S[SI_RH_CH4 * HX_TILE]);
        EMIT(OUT_NPP, mb.S[SI_X_NPP * HX_TILE]);
        EMIT(OUT_RH, mb.S[SI_X_RH * HX_TILE]);
        EMIT(OUT_GMST, gmst_out);
        EMIT(OUT_OCEAN_TAS, ocean_tas_out);
        EMIT(OUT_FLUX_MIXED, hf_mixed_out);
        EMIT(OUT_FLUX_INTERIOR, hf_int_out);
        EMIT(OUT_TIMESTEPS, (double)mb.timesteps);
        if (mb.X) {
          EMIT(OUT_UPTAKE_HL, mb.X[XS_UPTAKE_HL * HX_TILE]);
          EMIT(OUT_UPTAKE_LL, mb.X[XS_UPTAKE_LL * HX_TILE]);
          EMIT(OUT_RH_DET, mb.X[XS_RH_DET * HX_TILE]);
          EMIT(OUT_RH_SOIL, mb.X[XS_RH_SOIL * HX_TILE]);
        }
        }
        if (BIOMES && ALLOUT) { /* <biome>.<name>: the biome's own pools and final NPP / RH */
          const int bf[BO_COUNT] = {BF_VEG, BF_DET, BF_SOIL, BF_PERMAFROST, BF_THAWED, BF_X_NPP, BF_X_RH};
          for (int ib = 0; ib < C.n_biomes; ++ib)
            for (int k = 0; k < BO_COUNT; ++k) {
              const int slot_ = d.out_slot[OUT_COUNT + ib * BO_COUNT + k];
              if (slot_ >= 0)
                d.out[((size_t)slot_ * nyears_total + yi) * Mp + mo] = biome_of(mb, ib).f(bf[k]);
            }
        }
#undef EMIT
        }
      }
    }
#if HX_YEAR_SYNC
    /* r = the year a member stopped in (its barrier for that year is done), rend + 1 for a
     * member that ran the whole slab, base + 1 with no arrivals yet for a lane that never ran */
    for (; bars_done < (HX_YEAR_SYNC - 1) * (rend - base) + (rend - base + HX_YEAR_SYNC_EVERY - 1) / HX_YEAR_SYNC_EVERY; ++bars_done)
      year_barrier();
#endif
    if (TRACK) {
      /* hand the slab's record to the replay kernel: years a stopped member never reached carry
       * the last count forward; a record that overflowed fails the member */
      unsigned char *yc = d.YCNT + (size_t)s * d.ycnt_slab_stride + (size_t)tile * (HX_SLAB_YEARS * HX_BLOCK) + tid;
#pragma unroll
      for (int j = 0; j < HX_SLAB_YEARS; ++j) {
        if (j > 0 && ycnt[j] < ycnt[j - 1]) ycnt[j] = ycnt[j - 1];
        yc[j * HX_BLOCK] = entered ? ycnt[j] : 0;
      }
      if (entered && mb.trk_bad && mb.status == 0) {
        mb.status = HX_MEMBER_TRACKING;
        d.status[m] = mb.status;
        d.fail_year[m] = C.start_year + base + 1;
      }
    }
    if (lane_ok) {
      if (mb.status == 0) store_member(BS, mb);
      else ++failed;
    }
    if (d.slab_done) {
      /* streaming run: the slab's output rows leave for the host as soon as every tile has
       * finished it, so a failed member's NaNs are written here, not by a pass after the run */
      const int stn = d.status[m];
      if (stn > 0) nan_fill_rows(d, C.nrow - 1, mo, max(d.fail_year[m] - C.start_year - 1, base), rend);
    }
    if (S_IN_SMEM) { /* the state block goes back to global memory */
#pragma unroll 8
      for (int i = SI_REG_COUNT; i < SI_COUNT; ++i) gS[i * HX_BLOCK] = BS.S[i * HX_BLOCK];
    }
    if (bf_in_smem) {
      const int nf = C.n_biomes * BF_COUNT;
#pragma unroll 4
      for (int i = 0; i < nf; ++i) gBF[i * HX_BLOCK] = BS.BF[i * HX_BLOCK];
    }
    /* publish the tile's state: make this CTA's global stores visible, then release */
    __threadfence();
    __syncthreads(); /* also: everyone is done with slab[0] / row0 before they are refilled */
    if (tid == 0) {
      const unsigned done = (unsigned)s + 1u;
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(state + tile), "r"(done << 1) : "memory");
      if (d.slab_done) {
        /* the last tile to finish slab s raises the host's flag: a plain store into mapped
         * memory after a system-scope fence (the tiles are counted in device memory) */
        unsigned *slab_count = d.sched + 1 + ntiles;
        if (atomicAdd(slab_count + s, 1u) == (unsigned)ntiles - 1u) {
          __threadfence_system();
          *(volatile unsigned *)(d.slab_done + s) = 1u;
        }
      }
    }
  }
  flush_work(d, w, years_done, failed);
}

/* every map slot back to its initial state: pools and their CarbonAdditions are {own name: 1}
 * (fluxpool::set), the ocean's atmosphere copy is the atmosphere's, scratch slots are empty */
__global__ void hx_track_init_kernel(const __grid_constant__ HxDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.Mpad) return;
  const size_t tile = (size_t)(m / HX_BLOCK), ln = (size_t)(m % HX_BLOCK);
  double *T = d.T + tile * (size_t)(TS_COUNT * HX_NSRC) * HX_BLOCK + ln;
  uint32_t *K = d.TK + tile * (size_t)TS_COUNT * HX_BLOCK + ln;
  for (int i = 0; i < TS_COUNT; ++i) {
    int self = -1;
    if (i < HX_NPOOL) self = i;
    else if (i == TS_ATM_CPOOL) self = TS_ATMOS;
    else if (i >= TS_ADD_HL) self = TS_HL + (i - TS_ADD_HL);
    for (int s = 0; s < HX_NSRC; ++s) T[((size_t)i * HX_NSRC + s) * HX_BLOCK] = (s == self) ? 1.0 : 0.0;
    K[(size_t)i * HX_BLOCK] = self >= 0 ? (1u << self) : 0u;
  }
}

/* Replay of one slab's stash records into the members' source maps: one thread per (member,
 * two sources), six lanes per member, five members per warp so that a member never straddles
 * one.  The six lanes stage the member's record through shared memory: while stash st is mixed,
 * the (a, b) pairs of stash st+1 are already in flight as whole 128-byte lines. */
#ifndef HX_TRK_NS
#define HX_TRK_NS 2 /* sources per replay thread: 2-way ILP in every mix (150 vs 163 ms with 1) */
#endif
/* lanes per member: the twelve sources at HX_TRK_NS per lane -- six lanes, five members per
 * warp (two lanes of every warp idle; eight lanes per member with two of them mixing zeros kept
 * members aligned to the warp but wasted a quarter of every instruction) */
#define HX_TRK_LANES ((HX_NSRC + HX_TRK_NS - 1) / HX_TRK_NS)
#define HX_TRK_PER_WARP (32 / HX_TRK_LANES)      /* members per warp */
#define HX_TRK_MEMBERS (4 * HX_TRK_PER_WARP)     /* members per 128-thread CTA */
/* + 2: a member's two buffers start 300 words apart, so the five members of a warp read their
 * rows (one address per member) from different banks */
#define HX_TRK_SH_STRIDE (2 * HX_REC_MIX * HX_REC_ROW + 2) /* doubles of staging per member */
#define HX_TRK_SH_BYTES (HX_TRK_MEMBERS * HX_TRK_SH_STRIDE * 8)
/* 1 / t for an ordinary positive t (the caller has checked 1e-290 < t < 1e290): the hardware's
 * 20-bit seed and one cubic Newton step, relative error about 2 ulp -- the weight v of a mix
 * needs no more (tm_mix, hx_model.cuh), and the division operator's range test, its branch and
 * its two further steps were a third of the staging code */
__device__ __forceinline__ double rcp_ordinary(double t) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
  double e = fma(-t, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}
struct StagedRecord {
  const double2 *next; /* this lane's first pair of the stash to fetch next: the slab's record holds
                          pair (stash, k) of the member at [(stash * HX_REC_MIX + k) * stride] */
  size_t stride;       /* members per row */
  double *sh;          /* this member's two staging buffers, [2][HX_REC_MIX * HX_REC_ROW] */
  int lane, nst;
  unsigned mask;       /* the member's lanes within the warp */
  double2 v[(HX_REC_MIX + HX_TRK_LANES - 1) / HX_TRK_LANES]; /* this lane's pairs of the next stash, in flight */
  int staged;          /* stash whose pairs are in v */
  bool slow;           /* the stash just staged needs the general mix (see stash()) */
  __device__ __forceinline__ void begin(const double2 *rec) {
    next = rec + (size_t)lane * stride;
    staged = -1;
    slow = true;
    prefetch();
  }
  __device__ __forceinline__ void prefetch() {
    if (++staged >= nst) return;
    const double2 *q = next;
    const size_t step = (size_t)HX_TRK_LANES * stride;
#pragma unroll
    for (int j = 0; j < (HX_REC_MIX + HX_TRK_LANES - 1) / HX_TRK_LANES; ++j) {
      v[j] = (lane + HX_TRK_LANES * j < HX_REC_MIX) ? __ldcs(q) : make_double2(0.0, 0.0);
      q += step;
    }
    next += (size_t)HX_REC_MIX * stride;
  }
  /* rows (v = b / (a + b), a + b) of stash st in shared memory; st advances by one per call.
   * `slow` is set for the member's lanes when the stash holds anything the straight-line mix
   * does not cover: a total that is not an ordinary positive number -- other than the two that
   * are zero for decades and keep their test (R_T0, thawed permafrost before the first thaw;
   * R_OA, the ocean -> air flux while both surface boxes take carbon up) --, or a dump into the
   * deep ocean (a present R_DUMPx row) */
  __device__ __forceinline__ const double *stash(int st) {
    double *buf = sh + (st & 1) * (HX_REC_MIX * HX_REC_ROW);
    bool odd = false;
#pragma unroll
    for (int j = 0; j < (HX_REC_MIX + HX_TRK_LANES - 1) / HX_TRK_LANES; ++j) {
      const int k = lane + HX_TRK_LANES * j;
      if (k < HX_REC_MIX) {
        const double total = __dadd_rn(v[j].x, v[j].y);
        const bool dump = (k == R_DUMP0 || k == R_DUMP1); /* absent dumps carry a NaN flux */
        double w;
        if (__builtin_expect(total > 1e-290 && total < 1e290, 1)) {
          w = __dmul_rn(v[j].y, rcp_ordinary(total));
          odd = odd || dump;
        } else {
          /* b (1 / total), not b / total: a zero flux would send the division through its slow
           * path */
          w = (total != 0.0 && total == total) ? __dmul_rn(v[j].y, 1.0 / total) : 0.0;
          odd = odd || (dump ? (total == total) : !(total == 0.0 && (k == R_T0 || k == R_OA)));
        }
        *reinterpret_cast<double2 *>(buf + k * HX_REC_ROW) = make_double2(w, total);
      }
    }
    slow = __any_sync(mask, odd);
    prefetch();
    __syncwarp(mask);
    return buf;
  }
};

__global__ void __launch_bounds__(128, HX_TRK_NS == 1 ? 5 : HX_TRK_NS == 2 ? 3 : 2)
hx_track_kernel(const __grid_constant__ HxDev d, const __grid_constant__ HxConst C, int r0, int r1) {
  extern __shared__ __align__(16) double sh_dyn[];
  double (*sh)[HX_TRK_SH_STRIDE] = reinterpret_cast<double (*)[HX_TRK_SH_STRIDE]>(sh_dyn);
  const int wl = threadIdx.x & 31, wm = wl / HX_TRK_LANES; /* lane and member within the warp */
  if (wm >= HX_TRK_PER_WARP) return;                      /* the warp's spare lanes */
  const int cm = (threadIdx.x >> 5) * HX_TRK_PER_WARP + wm; /* member within the CTA */
  const int m = blockIdx.x * HX_TRK_MEMBERS + cm, s = wl % HX_TRK_LANES;
  if (m >= d.Mpad) return;
  if (d.status[m] < 0) return; /* padding member (all its lanes leave together) */
  const size_t tile = (size_t)(m / HX_BLOCK), ln = (size_t)(m % HX_BLOCK);
  double *T = d.T + tile * (size_t)(TS_COUNT * HX_NSRC) * HX_BLOCK + ln;
  uint32_t *TK = d.TK + tile * (size_t)TS_COUNT * HX_BLOCK + ln;
  const unsigned char *yc = d.YCNT + tile * (size_t)(HX_SLAB_YEARS * HX_BLOCK) + ln;
  const int nyears = r1 - r0;
  StagedRecord fetch;
  fetch.stride = (size_t)d.Mpad;
  fetch.sh = &sh[cm][0];
  fetch.lane = s;
  fetch.nst = yc[(nyears - 1) * HX_BLOCK];
  fetch.mask = ((1u << HX_TRK_LANES) - 1u) << (wm * HX_TRK_LANES);
  fetch.begin(reinterpret_cast<const double2 *>(d.REC) + m);
  const bool good = track_replay<HX_TRK_NS>(T, TK, fetch, yc, HX_BLOCK, nyears,
                                            C.start_year + r0 + 1, s * HX_TRK_NS,
                                            (s + 1) * HX_TRK_NS, C.tracking_date, C.track_every, C.track_nrec,
                                    C.end_year, d.TO + m, d.TOK + m, (size_t)d.Mpad);
  /* a bad mix is reported through the replay's own word: the status words belong to the run
   * kernel, which may already be computing the next slab on the other stream.  Replays of one
   * engine run in order, so the first failing slab's year sticks; the lanes of a member that
   * disagree all store the same value. */
  if (s * HX_TRK_NS < HX_NSRC && !good && d.trk_fail[m] == 0) d.trk_fail[m] = C.start_year + r0 + 1;
}

__global__ void hx_track_merge_kernel(const __grid_constant__ HxDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.Mpad) return;
  const int y = d.trk_fail[m];
  if (y == 0 || d.status[m] < 0) return;
  if (d.status[m] == 0 || d.fail_year[m] > y) {
    d.status[m] = HX_MEMBER_TRACKING;
    d.fail_year[m] = y;
  }
}

/* failed members report NaN from the failing year on (the reference stops producing output) */
__global__ void hx_nan_fill_kernel(const __grid_constant__ HxDev d, int start_year, int nyears,
                                   int nsel, int yr0, int yr1) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.Mpad) return;
  const int st = d.status[m];
  if (st <= 0) return;
  int first = d.fail_year[m] - start_year - 1; /* output index of the failing year */
  if (first < yr0) first = yr0;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const int col = d.api_of_dev[m];
  for (int s = 0; s < nsel; ++s)
    for (int yi = first; yi < yr1; ++yi) d.out[((size_t)s * nyears + yi) * d.Mpad + col] = nan;
}

/* ---- host-callable launchers ---- */
cudaError_t launch_setup(const HxDev &d, const HxConst &C, cudaStream_t st, int phase) {
  hx_setup_kernel<<<d.Mpad / HX_BLOCK, HX_BLOCK, 0, st>>>(d, C, phase);
  return cudaGetLastError();
}
cudaError_t launch_spinup(const HxDev &d, const HxConst &C, cudaStream_t st) {
  if (d.BF) hx_spinup_kernel<true><<<d.Mpad / HX_BLOCK, HX_BLOCK, 0, st>>>(d, C, 0, -1);
  else hx_spinup_kernel<false><<<d.Mpad / HX_BLOCK, HX_BLOCK, 0, st>>>(d, C, 0, -1);
  return cudaGetLastError();
}
cudaError_t launch_spinup_one(const HxDev &d, const HxConst &C, int member, cudaStream_t st) {
  hx_spinup_kernel<false><<<1, HX_BLOCK, 0, st>>>(d, C, member - member % HX_BLOCK, member);
  return cudaGetLastError();
}
template <bool TRACK, bool CONSTR, int MINCTAS, bool ALLOUT = true, bool BIOMES = false,
          bool NBP = CONSTR, bool EXACT = false, bool GAS = false, bool LAT = false, bool RF4 = false>
static cudaError_t launch_run_t(const HxDev &d, const HxConst &C, int r0, int r1, cudaStream_t st) {
  /* the function attribute and the occupancy are per device (context): one cache slot per
   * device ordinal, so that engines on several GPUs can live in one process */
  static int resident_of[HX_MAX_DEVICES] = {};
  int dev = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 != cudaSuccess) return e0;
  if (dev < 0 || dev >= HX_MAX_DEVICES) return cudaErrorInvalidDevice;
  if (!resident_of[dev]) {
    cudaError_t e = cudaFuncSetAttribute(hx_run_kernel<TRACK, CONSTR, MINCTAS, ALLOUT, BIOMES, NBP, EXACT, GAS, LAT, RF4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)run_smem_bytes<MINCTAS, LAT>());
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hx_run_kernel<TRACK, CONSTR, MINCTAS, ALLOUT, BIOMES, NBP, EXACT, GAS, LAT, RF4>,
                                                      HX_BLOCK, run_smem_bytes<MINCTAS, LAT>());
    if (e != cudaSuccess) return e;
    resident_of[dev] = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int resident = resident_of[dev];
  /* persistent CTAs: never more than can be co-resident (an item may wait on another CTA) */
  const int ntiles = d.Mpad / HX_BLOCK;
  const int grid = ntiles < resident ? ntiles : resident;
  const int nslab = (r1 - r0 + HX_SLAB_YEARS - 1) / HX_SLAB_YEARS;
  cudaError_t e = cudaMemsetAsync(d.sched, 0, (size_t)(ntiles + 1 + nslab) * sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  hx_run_kernel<TRACK, CONSTR, MINCTAS, ALLOUT, BIOMES, NBP, EXACT, GAS, LAT, RF4><<<grid, HX_BLOCK, run_smem_bytes<MINCTAS, LAT>(), st>>>(d, C, r0, r1);
  return cudaGetLastError();
}
cudaError_t launch_run(const HxDev &d, const HxConst &C, int r0, int r1, cudaStream_t st) {
  /* Two register budgets.  Big ensembles run the 168-register build: 3 CTAs (12 warps) per SM
   * hide more FP64 latency than its few spills cost (39.1 vs 40.8 ms at 65 536 members).  When
   * every tile is resident anyway at 2 CTAs per SM the run is pure latency and the spill-free
   * 230-register build wins (17.0 vs 20.5 ms at 1 024 members).  The tracking build is bound by
   * its map traffic, not by occupancy, and spills badly at 168 registers. */
  /* biome-split pools: general builds (constraints and every output; no tracking), with and
   * without the NBP machinery */
  const bool exact = (C.flags & HX_FLAG_EXACT_ATTEMPTS) != 0;
  if (d.BF) {
    if (!d.GP && exact && d.constrained <= 1) /* abandoned ODE attempts executed (the NBP build always does) */
      return d.constrained ? launch_run_t<false, true, 2, true, true, false, true>(d, C, r0, r1, st)
                           : launch_run_t<false, false, 2, true, true, false, true>(d, C, r0, r1, st);
    if (d.GP) { /* ... with per-member N2O / halocarbon parameters */
      if (d.constrained > 1) return launch_run_t<false, true, 2, true, true, true, false, true>(d, C, r0, r1, st);
      if (exact)
        return d.constrained ? launch_run_t<false, true, 2, true, true, false, true, true>(d, C, r0, r1, st)
                             : launch_run_t<false, false, 2, true, true, false, true, true>(d, C, r0, r1, st);
      if (d.constrained) return launch_run_t<false, true, 2, true, true, false, false, true>(d, C, r0, r1, st);
      return launch_run_t<false, false, 2, true, true, false, false, true>(d, C, r0, r1, st);
    }
    if (d.constrained > 1) return launch_run_t<false, true, 2, true, true, true>(d, C, r0, r1, st);
    if (d.constrained) return launch_run_t<false, true, 2, true, true, false>(d, C, r0, r1, st);
    /* no constraint, no lo_warming_ratio: the biome loops without the constraint machinery;
     * CO2 / Tgav only: without the other outputs' code either */
    return d.out_minimal == 1 ? launch_run_t<false, false, 2, false, true, false>(d, C, r0, r1, st)
                         : launch_run_t<false, false, 2, true, true, false>(d, C, r0, r1, st);
  }
  if (d.T) { /* carbon tracking: the record-only builds */
    if (d.GP) { /* ... with per-member N2O / halocarbon parameters */
      if (d.constrained) return launch_run_t<true, true, HX_TRACK_CTAS, true, false, true, false, true>(d, C, r0, r1, st);
      return exact ? launch_run_t<true, false, HX_TRACK_CTAS, true, false, false, true, true>(d, C, r0, r1, st)
                   : launch_run_t<true, false, HX_TRACK_CTAS, true, false, false, false, true>(d, C, r0, r1, st);
    }
    if (d.constrained) return launch_run_t<true, true, HX_TRACK_CTAS>(d, C, r0, r1, st);
    if (exact) return launch_run_t<true, false, HX_TRACK_CTAS, true, false, false, true>(d, C, r0, r1, st);
    return d.out_minimal == 1 ? launch_run_t<true, false, HX_TRACK_CTAS, false>(d, C, r0, r1, st)
                         : launch_run_t<true, false, HX_TRACK_CTAS>(d, C, r0, r1, st);
  }
  if (d.GP) { /* per-member N2O / halocarbon parameters: the GAS builds (every output) */
    if (d.constrained > 1) return launch_run_t<false, true, 2, true, false, true, false, true>(d, C, r0, r1, st);
    if (exact)
      return d.constrained ? launch_run_t<false, true, 2, true, false, false, true, true>(d, C, r0, r1, st)
                           : launch_run_t<false, false, 2, true, false, false, true, true>(d, C, r0, r1, st);
    if (d.constrained) return launch_run_t<false, true, 2, true, false, false, false, true>(d, C, r0, r1, st);
    return launch_run_t<false, false, 2, true, false, false, false, true>(d, C, r0, r1, st);
  }
  if (exact) { /* the builds that execute abandoned ODE attempts */
    if (d.constrained > 1) return launch_run_t<false, true, 2>(d, C, r0, r1, st); /* NBP: always exact */
    return d.constrained ? launch_run_t<false, true, 2, true, false, false, true>(d, C, r0, r1, st)
                         : launch_run_t<false, false, 2, true, false, false, true>(d, C, r0, r1, st);
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool small = d.Mpad / HX_BLOCK <= 2 * sms;
  if (d.constrained > 1) /* an NBP constraint somewhere */
    return small ? launch_run_t<false, true, 2>(d, C, r0, r1, st)
                 : launch_run_t<false, true, HX_RUN_MIN_CTAS>(d, C, r0, r1, st);
  if (d.constrained) {   /* the other constraints, lo_warming_ratio: no NBP machinery */
    if (d.out_minimal == 1)
      return small ? launch_run_t<false, true, 2, false, false, false>(d, C, r0, r1, st)
                   : launch_run_t<false, true, HX_RUN_MIN_CTAS, false, false, false>(d, C, r0, r1, st);
    return small ? launch_run_t<false, true, 2, true, false, false>(d, C, r0, r1, st)
                 : launch_run_t<false, true, HX_RUN_MIN_CTAS, true, false, false>(d, C, r0, r1, st);
  }
  /* "small": at most one CTA per SM, i.e. every warp has a scheduler to itself -> the LAT build */
  /* (HX_NO_LAT=1 keeps small ensembles on the general build: sanitizer runs of that build) */
  static const bool no_lat = std::getenv("HX_NO_LAT") != nullptr;
  const bool lone = !no_lat && d.Mpad / HX_BLOCK <= sms;
  if (d.out_minimal == 1)
    return lone ? launch_run_t<false, false, 2, false, false, false, false, false, true>(d, C, r0, r1, st)
                : launch_run_t<false, false, HX_RUN_MIN_CTAS, false>(d, C, r0, r1, st);
  if (d.out_minimal == 2) /* R's default four: the CO2 / Tgav-only builds plus RF_tot and RF_CO2 */
    return lone ? launch_run_t<false, false, 2, false, false, false, false, false, true, true>(d, C, r0, r1, st)
                : launch_run_t<false, false, HX_RUN_MIN_CTAS, false, false, false, false, false, false, true>(d, C, r0, r1, st);
  return lone ? launch_run_t<false, false, 2, true, false, false, false, false, true>(d, C, r0, r1, st)
              : launch_run_t<false, false, HX_RUN_MIN_CTAS>(d, C, r0, r1, st);
}
size_t track_record_bytes_per_cta() {
  return (size_t)HX_REC_STASH_MAX * HX_REC_N * HX_BLOCK * sizeof(double);
}
size_t track_ycnt_bytes_per_tile() { return (size_t)HX_SLAB_YEARS * HX_BLOCK; }
int track_slab_years() { return HX_SLAB_YEARS; }
cudaError_t launch_track(const HxDev &d, const HxConst &C, int r0, int r1, cudaStream_t st) {
  if (HX_TRK_SH_BYTES > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hx_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HX_TRK_SH_BYTES);
    if (e != cudaSuccess) return e;
  }
  hx_track_kernel<<<(unsigned)((d.Mpad + HX_TRK_MEMBERS - 1) / HX_TRK_MEMBERS), 128, HX_TRK_SH_BYTES, st>>>(d, C, r0, r1);
  return cudaGetLastError();
}
cudaError_t launch_track_merge(const HxDev &d, cudaStream_t st) {
  hx_track_merge_kernel<<<(d.Mpad + 255) / 256, 256, 0, st>>>(d);
  return cudaGetLastError();
}
cudaError_t launch_track_init(const HxDev &d, cudaStream_t st) {
  hx_track_init_kernel<<<(d.Mpad + 255) / 256, 256, 0, st>>>(d);
  return cudaGetLastError();
}
cudaError_t launch_nan_fill(const HxDev &d, const HxConst &C, int nsel, int yr0, int yr1,
                            cudaStream_t st) {
  hx_nan_fill_kernel<<<(d.Mpad + 255) / 256, 256, 0, st>>>(d, C.start_year, C.nrow - 1, nsel, yr0,
                                                           yr1);
  return cudaGetLastError();
}

/* diagnostics: the interleaved transcendentals (hx_exp_n / hx_log_n, hx_model.cuh) next to the
 * library routines they restate, value by value (tests compare the two bit for bit) */
__global__ void hx_transc_check_kernel(const double *x, double *fast, double *lib, int n, int which) {
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  double a[4], y[4];
  for (int k = 0; k < 4; ++k) a[k] = x[min(i0 + k, n - 1)];
  if (which == 0) hx_exp_n<4, false>(a, y);
  else if (which == 1) hx_exp_n<4, true>(a, y);
  else if (which == 2) hx_log_n<4>(a, y);
  else { /* hx_div: x holds (numerator, denominator) pairs, outputs in the numerator's slot */
    y[0] = hx_div(a[0], a[1]); y[1] = 0.0;
    y[2] = hx_div(a[2], a[3]); y[3] = 0.0;
  }
  for (int k = 0; k < 4 && i0 + k < n; ++k) {
    fast[i0 + k] = y[k];
    lib[i0 + k] = which == 0 ? exp(a[k]) : which == 1 ? exp10(a[k]) : which == 2 ? log(a[k])
                : (k & 1) ? 0.0 : a[k] / a[min(k + 1, 3)];
  }
}

} // namespace hx

extern "C" int hx_diag_transcendentals(int32_t device, int32_t which, const double *x, double *fast,
                                       double *lib, int32_t n) {
  if (!x || !fast || !lib || n <= 0 || which < 0 || which > 3) return HX_ERR_ARG;
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return HX_ERR_CUDA;
  double *dx = nullptr, *df = nullptr, *dl = nullptr;
  const size_t bytes = (size_t)n * sizeof(double);
  cudaError_t e = cudaMalloc(&dx, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&df, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&dl, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int threads = (n + 3) / 4;
    hx::hx_transc_check_kernel<<<(threads + 127) / 128, 128>>>(dx, df, dl, n, which);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(fast, df, bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(lib, dl, bytes, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(df); cudaFree(dl);
  cudaSetDevice(prev);
  return e == cudaSuccess ? HX_OK : HX_ERR_CUDA;
}
