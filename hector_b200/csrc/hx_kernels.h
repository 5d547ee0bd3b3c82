/* hx_kernels.h -- device pointer bundle + kernel launchers (host <-> device boundary inside
 * the library). */
#ifndef HX_KERNELS_H
#define HX_KERNELS_H

#include <cuda_runtime.h>

#include "hx_layout.h"

#define HX_BLOCK 128      /* threads (= members) per CTA; one scenario per CTA */
#ifndef HX_CONV_UNROLL
#define HX_CONV_UNROLL 8 /* history rows per trip of the slab prepass of the DOECLIM convolution */
#endif
#ifndef HX_TRACK_CTAS
#define HX_TRACK_CTAS 2 /* resident CTAs per SM of the record-only (tracking) run kernel */
#endif
#ifndef HX_CONV_L2_AHEAD
#define HX_CONV_L2_AHEAD 3 /* trips ahead of the prepass whose history rows are prefetched into L2 (0 = off) */
#endif
#ifndef HX_FORC_AHEAD
#define HX_FORC_AHEAD 1 /* the forcing's and DOECLIM's constants are requested as soon as the solver is done */
#endif
#ifndef HX_OH_AHEAD
#define HX_OH_AHEAD 2 /* the OH / CH4 constants are requested ahead: 1 = before the year barrier, 2 = after the solver of the year before */
#endif
#ifndef HX_YEAR_SYNC_EVERY
#define HX_YEAR_SYNC_EVERY 1 /* the top-of-year barrier every n-th year of a work item */
#endif
#ifndef HX_YEAR_SYNC
#define HX_YEAR_SYNC 1 /* one CTA barrier per simulated year: the warps share instruction fetches */
#endif
#ifndef HX_RUN_MIN_CTAS
#define HX_RUN_MIN_CTAS 2 /* resident CTAs per SM the run kernel is compiled for.  2: 246 registers, no
                             spills, and room in shared memory for the tile's whole state block
                             (31.6 ms at 65 536 members); 3: 168 registers with spills, state in
                             global memory / L1 (32.5 ms); 2 without the state block: 32.7 ms */
#endif

struct HxDev {
  int32_t Mpad;             /* padded member count, multiple of HX_BLOCK */
  const double *P;          /* [tile][PI_COUNT | DI_COUNT][128]: parameters, then derived constants */
  double *S;                /* [SI_COUNT][Mpad] */
  double *D;                /* tile 0's derived constants inside P (kernels derive it from P) */
  double *ker;              /* [HX_KER_ROWS(nrow)][Mpad]  DOECLIM lag kernel K(j), zero padded */
  double *conv;             /* [HX_SLAB_YEARS][Mpad] per-slab partial convolution sums */
  double *sst_hist;         /* [nrow][Mpad] */
  double *tland_hist;       /* [nrow][Mpad] */
  double *out;              /* [nsel][nrow-1][Mpad], columns in API member order */
  double *X;                /* [tile][XS_COUNT][128] scratch rows of the per-stash outputs, or null */
  const int32_t *api_of_dev; /* [Mpad] API index (= output column) of each device member, -1 = padding */
  const double *scen;       /* [n_scen][nrow][SC_STRIDE] */
  const int32_t *block_scen; /* [Mpad / HX_BLOCK] scenario of each CTA */
  int32_t *status;          /* [Mpad]; -1 = padding lane */
  int32_t *fail_year;       /* [Mpad] */
  int32_t *spinup_steps;    /* [Mpad] */
  unsigned long long *counters; /* [HX_NCOUNTERS] */
  unsigned *sched;              /* [1 + tiles + slabs]: (unused), per-tile (progress << 1 | busy), tiles done per slab */
  int32_t out_slot[HX_OUT_IDS]; /* output id -> slot in `out`, -1 = not recorded; ids from
                                   OUT_COUNT on are the per-biome outputs */
  int32_t constrained;      /* 0 none; 1 some scenario carries a CO2 / CH4 / RF_tot / tas
                               constraint or a member a lo_warming_ratio; 2 an NBP constraint */
  int32_t out_minimal;      /* 1: only CO2_concentration and/or global_tas are recorded; 2: only those
                               and RF_tot / RF_CO2 (R's default four); 0: anything else */
  int32_t n_out;            /* recorded outputs (slots of `out`) */
  /* hx_run_stream: [slabs of the launch] in mapped host memory, set to 1 when every tile has
   * finished the slab -- the host copies a slab's output rows out while later slabs still run
   * (null: off) */
  unsigned *slab_done;
  /* carbon tracking (null unless a tracking date was set) */
  double *T;                /* [tile][TS_COUNT * HX_NSRC][128] source fractions */
  uint32_t *TK;             /* [tile][TS_COUNT][128] key masks */
  double *REC;              /* [HX_REC_STASH_MAX][HX_REC_MIX][member] (a, b) pairs: stash records of the slab
                               being run (hx_model.cuh, "Carbon tracking") */
  unsigned char *YCNT;      /* [tile][HX_SLAB_YEARS][128] stashes recorded up to each year's end */
  /* a tracked launch may cover several slabs: slab s of the launch records into REC + s *
   * rec_slab_stride (doubles) and YCNT + s * ycnt_slab_stride (bytes) */
  size_t rec_slab_stride, ycnt_slab_stride;
  double *TO;               /* [track_nrec][HX_NPOOL * HX_NSRC][Mpad] recorded fractions */
  /* biomes (null with the single global biome) */
  const double *BP;         /* [tile][n_biomes * BP_COUNT][128] per-biome parameters */
  double *BF;               /* [tile][n_biomes * BF_COUNT][128] per-biome pools and factors */
  uint32_t *TOK;            /* [track_nrec][HX_NPOOL][Mpad] recorded key masks */
  /* per-member N2O / halocarbon parameters (null: host series per scenario, the default) */
  const double *GP;         /* [tile][GP_COUNT][128] */
  double *GF;               /* [tile][GF_COUNT][128] */
  const double *scen_gas;   /* [n_scen][nrow][HX_GAS_COLS] emissions */
  int32_t *trk_fail;        /* [Mpad] first year whose replay saw a bad mix, 0 = none: written by
                               the replay kernel only, folded into status by hx_track_merge */
};

namespace hx {
/* phase: 1 = ocean rates + initial state (what the spin-up reads), 2 = DOECLIM matrices and lag
 * kernel, 3 = both */
cudaError_t launch_setup(const HxDev &d, const HxConst &C, cudaStream_t st, int phase = 3);
cudaError_t launch_spinup(const HxDev &d, const HxConst &C, cudaStream_t st);
cudaError_t launch_spinup_one(const HxDev &d, const HxConst &C, int member, cudaStream_t st);
cudaError_t launch_run(const HxDev &d, const HxConst &C, int r0, int r1, cudaStream_t st);
cudaError_t launch_track_init(const HxDev &d, cudaStream_t st);
size_t track_record_bytes_per_cta(); /* size of one tile's REC block */
size_t track_ycnt_bytes_per_tile();
int track_slab_years();
/* replay the records of rows r0+1 .. r1 (one slab, as just run) into the source maps */
cudaError_t launch_track(const HxDev &d, const HxConst &C, int r0, int r1, cudaStream_t st);
/* fold the replays' failure words into status / fail_year (earliest failing year wins) */
cudaError_t launch_track_merge(const HxDev &d, cudaStream_t st);
cudaError_t launch_nan_fill(const HxDev &d, const HxConst &C, int nsel, int yr0, int yr1,
                            cudaStream_t st);
}
#endif
