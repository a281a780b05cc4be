/* surfdisp_b200.h -- C ABI of the B200-native batched surface-wave dispersion forward solver.
 *
 * Drop-in boundary for the hot path of 001cat/pySurfInv (citations relative to the reference tree):
 *   - fast_surf_()            replaces the gfortran symbol behind the f2py module
 *                             (fast_surf_src/fast_surf.f:2-5, signature fast_surf_src/fast_surf.pyf:6-19,
 *                              called from models.py:27 and senskernel.py:188)
 *   - surfdisp_batch()        the same computation for M models at once on device buffers
 *                             (replaces M sequential FAST_SURF calls of point.py:19 / models.py:115-121)
 *   - surfdisp_misfit_batch() replaces Point.misfit (point.py:15-31) / PointCascadia.misfit
 *                             (point.py:337-366) for M models
 *   - surfdisp_host_batch()   host-buffer convenience wrapper (H2D, solve, D2H) -- what a ctypes/cffi
 *                             caller without device memory management uses
 *
 * All entry points are re-entrant (no global mutable state), never print, never abort the process.
 * Pointers documented "device" must be device-accessible; everything else is host memory.
 * Return value: 0 on success, negative SURFDISP_E* on error.
 */
#ifndef SURFDISP_B200_H
#define SURFDISP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURFDISP_KIND_LOVE 1      /* fast_surf kind0 = 1 */
#define SURFDISP_KIND_RAYLEIGH 2  /* fast_surf kind0 = 2 */

#define SURFDISP_MAX_PERIODS 200  /* fast_surf.f:9  (nper) */
#define SURFDISP_MAX_LAYERS 1000  /* fast_surf.f:9  (nsize) */

#define SURFDISP_EINVAL (-1)      /* bad argument */
#define SURFDISP_ENOMEM (-2)      /* workspace too small / allocation failed */
#define SURFDISP_ECUDA (-3)       /* CUDA runtime error (see surfdisp_last_cuda_error) */

/* per-model status bits written to flags[] */
#define SURFDISP_F_NO_ROOT_FIRST 1   /* no root at the first period (calcul.f:203-212): nfound = 0 */
#define SURFDISP_F_NO_ROOT_AT_K 2    /* scan failed at a later period (calcul.f:218-219): nfound = k-1 */
#define SURFDISP_F_ROOT_ABOVE_HS 4   /* polished root above the half-space velocity (calcul.f:191) */
#define SURFDISP_F_SCAN_LIMIT 8      /* scan hit the iteration guard (non-finite secular function) */
#define SURFDISP_F_LSTOP 16          /* sequential polish did not converge in 50 cycles: the reference aborts
                                        the call there (surfa.f:17-28, calcul.f:173-189); nfound = 0 */

/* Solver options; defaults are the constants hard-coded in the reference (init.f:25,43-58). */
typedef struct SurfdispOpts {
  float dc;                /* scan step, km/s                 (init.f:25   dc=0.01)  */
  float fact;              /* layer-drop depth in wavelengths (init.f:25   fact=4)   */
  float t_base;            /* reference period of Q, s        (fast_surf.f:77  t_base=1) */
  int ndiv;                /* sub-layers per layer for the energy integrals (init.f:25 ndiv=5);
                              per model it is clamped to cap/(n-1) like a fresh process would
                              (surfa.f:783-784, 414-415) */
  int ndiv_cap_rayleigh;   /* 99  (surfa.f:783) */
  int ndiv_cap_love;       /* 999 (surfa.f:414) */
  int atten;               /* KEY_ATTEN (init.f:43), 1 */
  int flatten;             /* earth flattening (calcul.f:133), 1 */
  int stale_mmax;          /* 1: period k refreshes only the layers kept by period k-1, as the
                              reference does (calcul.f:112-133); 0: refresh all layers */
  int compute_group;       /* 1: also group velocity (REIGEN/LEIGEN); 0: phase velocity only */
  int exact_scan;          /* 0 (default): after the first period the root is bracketed by trial velocities clustered
                              around the extrapolation of the previous roots (guard points at c1 and half way
                              rule out an odd number of skipped roots) and taken by inverse interpolation;
                              1: every period evaluates every grid point c1 + i*dc like calcul.f:155-167 and
                              polishes by uniform section (slower; differs only where two roots of different
                              modes lie between c1 and the tracked root) */
  int group_f64;           /* Rayleigh group velocity (REIGEN, surfa.f:714-1190): precision of the ODE state and of the
                              energy sums.  0 (default): float32, the two half-space solutions re-orthogonalised after
                              every sub-layer; 1: float64 like the reference (implicit double precision), re-orthogonalised
                              every 8 layers.  The two agree to 6e-6 km/s (2e-5 on 500-layer stacks) and have the same
                              error statistics against the reference: the noise of U comes from the float32 root c,
                              not from the integration */
} SurfdispOpts;

void surfdisp_default_opts(SurfdispOpts* o);

/* Bytes of device scratch surfdisp_batch needs for this problem size. */
size_t surfdisp_workspace_bytes(int n_models, int n_layers_max, int n_periods);

/* Batched forward solve on device buffers.
 *   kind          SURFDISP_KIND_LOVE / SURFDISP_KIND_RAYLEIGH
 *   n_layers      device int[M]: layers per model incl. the half-space (2 <= n <= n_layers_max)
 *   layers        device float[5][M][n_layers_max] in fast_surf argument order
 *                 (a = Vp, b = Vs, rho, d = thickness km, qs = 1/Qs); layer 0 is the top, the last
 *                 layer is the half-space (its thickness is ignored); b = 0 marks a water top layer
 *   periods       HOST float[K], seconds, shared by all models, K <= SURFDISP_MAX_PERIODS
 *   c_out, u_out  device float[M][K] phase / group velocity; zero beyond nfound[m]. u_out may be NULL
 *   nfound        device int[M]   = reference imax(1), the number of periods with a root
 *   flags         device int[M] or NULL, SURFDISP_F_* bits
 *   workspace     device scratch of at least surfdisp_workspace_bytes()
 *   stream        cudaStream_t (as void*), NULL = default stream.  The call is asynchronous.
 */
int surfdisp_batch(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max,
                   const int* n_layers, const float* layers, int n_periods, const float* periods,
                   float* c_out, float* u_out, int* nfound, int* flags, void* workspace,
                   size_t workspace_bytes, void* stream);

/* surfdisp_batch with a NEIGHBOUR CURVE per model: c_hint device float[M][K] (or NULL) = phase velocities of a nearby
 * model on the same periods -- in a Monte-Carlo walk the chain's current model, of which the proposal is a small
 * perturbation.  From the second period on the root is searched around hint(k) + [c(k-1) - hint(k-1)] (the difference
 * to the neighbour's curve, extrapolated) instead of around the extrapolation of the model's own earlier roots; the
 * sign guards and the fall-back to the reference's scan are unchanged, so results do not depend on the hint (a wrong,
 * zero or missing hint costs sweeps).  Rows with zeros are "no hint". */
int surfdisp_batch_hinted(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                          const float* layers, int n_periods, const float* periods, const float* c_hint, float* c_out,
                          float* u_out, int* nfound, int* flags, void* workspace, size_t workspace_bytes, void* stream);

/* Rayleigh phase velocities AND their partial derivatives with respect to Vp, Vs and density of every layer of each
 * period's (attenuation-corrected, flattened) model: REIGEN's dcda, dcdb, dcdr (surfa.f:1130-1135, 1179-1185,
 * 1202-1208) -- computed by the reference along with the group velocity and kept in COMMON /derivd/; here they are an
 * output.  They replace the 2 n + 1 forward solves per model of the finite-difference SensKernelPert
 * (senskernel.py:130-158).  dcda / dcdb / dcdr: device float[M][K][n_layers_max] (zero below the layer-dropping depth,
 * in liquid layers, beyond nfound).  The other arguments as surfdisp_batch. */
int surfdisp_partials_batch(const SurfdispOpts* opts, int n_models, int n_layers_max, const int* n_layers,
                            const float* layers, int n_periods, const float* periods, float* c_out, float* dcda,
                            float* dcdb, float* dcdr, int* nfound, int* flags, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Per-model misfit of predicted phase velocities against one observed curve.
 *   mode 0: Point.misfit (point.py:15-31); mode 1: PointCascadia.misfit (point.py:337-366)
 *   c_pred   device float[M][K], nfound device int[M] (models with nfound < K get the failure
 *            sentinel (88888, 88888, 0), models.py:29-32 + point.py:20-21)
 *   obs, sigma  HOST float[K]; mask HOST unsigned char[K] or NULL (1 = observation used)
 *   periods  HOST float[K] (only read in mode 1 for the T <= 40 s split)
 *   out      device float[M][3] = (misfit, chiSqr, L)
 */
int surfdisp_misfit_batch(int mode, int n_models, int n_periods, const float* c_pred, const int* nfound,
                          const float* obs, const float* sigma, const unsigned char* mask,
                          const float* periods, float* out, void* stream);

/* Host-buffer wrapper: copies inputs to the device, solves, copies results back and synchronises (through
 * surfdisp_host_batch_pipelined, 8 chunks from 65536 models on).  All pointers are HOST memory (pinned memory
 * makes the copies asynchronous).  device = CUDA device ordinal.  The device block and the two streams belong
 * to a per-host-thread context that is kept between calls (the block only grows, up to 1 GiB kept): a caller
 * that loops over single models like models.py:27 does not pay cudaMalloc / cudaStreamCreate per call.
 * Re-entrant: nothing is shared between host threads.  surfdisp_host_release() frees the calling thread's
 * context (optional; it is also freed at thread exit). */
int surfdisp_host_batch(const SurfdispOpts* opts, int device, int kind, int n_models, int n_layers_max,
                        const int* n_layers, const float* layers, int n_periods, const float* periods,
                        float* c_out, float* u_out, int* nfound, int* flags);
void surfdisp_host_release(void);

/* Tuning: from how many models on the later periods of the root search run as two launches -- a fast-path launch
 * without scan code that hands the models whose period needs the point-by-point scan (calcul.f:155-167) over to the
 * general launch, which resumes them at that period (default 98304; n_models <= 0 restores it).  Results do not
 * depend on it (tests/test_gpu_parity.py: bit-identical).  Process-wide. */
void surfdisp_set_split_min_models(int n_models);

/* The same with caller-owned device memory and streams, as a pipeline: the batch is cut into n_chunks chunks;
 * chunk i is copied host->device on copy_stream while chunk i-1 is prepared and its first period searched on
 * compute_stream; the later periods run as one launch over the whole batch; the group velocities are computed
 * chunk by chunk, each chunk copied device->host under the next one.  Host pointers should be pinned (pageable
 * memory works but serialises the copies).  device_buffer: surfdisp_pipelined_bytes() bytes of device memory.
 * Synchronises both streams: the host outputs are valid on return.  Replaces the per-model host round trip of
 * models.py:27 (one f2py call per model) for a batch. */
size_t surfdisp_pipelined_bytes(int n_models, int n_layers_max, int n_periods);
int surfdisp_host_batch_pipelined(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max,
                                  const int* n_layers, const float* layers, int n_periods, const float* periods,
                                  float* c_out, float* u_out, int* nfound, int* flags, void* device_buffer,
                                  size_t device_bytes, int n_chunks, void* compute_stream, void* copy_stream);

/* ABI-level replacement of the gfortran symbol FAST_SURF (fast_surf.f:2-5): one model, all arguments
 * by reference, host memory.  cvper has 200 entries of which the first *ncvper are used (init.f:62-72).
 * Unlike the reference the four 200-long outputs are fully defined: zero beyond the
 * found prefix and for the wave type not requested. */
void fast_surf_(const int* n_layer0, const int* kind0, const float* a_ref0, const float* b_ref0,
                const float* rho_ref0, const float* d_ref0, const float* qs_ref0, const float* cvper,
                const int* ncvper, float* uR0, float* uL0, float* cR0, float* cL0);

/* Device-side counters of executed work for the last surfdisp_batch on this workspace (roofline
 * numerator): out[0] = secular layer-steps, out[1] = secular sweeps,
 * out[2] = group-velocity sub-layer integrations, out[3] = models processed.  Host pointer. */
int surfdisp_read_counters(const void* workspace, unsigned long long out[4], void* stream);

/* surfdisp_batch plus CUDA-event timing of its three kernels (prep, phase 1 = root search, phase 2 = group
 * velocity); synchronises the stream.  kernel_ms is a HOST float[3]. */
int surfdisp_batch_profiled(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max,
                            const int* n_layers, const float* layers, int n_periods, const float* periods,
                            float* c_out, float* u_out, int* nfound, int* flags, void* workspace,
                            size_t workspace_bytes, void* stream, float kernel_ms[3]);

/* Register-resident micro-benchmarks on the current device: out[0] = FP32 FMA TFLOP/s,
 * out[1] = FP64 FMA TFLOP/s, out[2] = MUFU.EX2 T-op/s.  Denominators of the FP-pipe roofline. */
int surfdisp_measure_peaks(double out[3]);


/* ---------------------------------------------------------------------------------------------------------
 * Device-side model builder: free parameters of a Monte-Carlo walk -> layer stacks in the layout
 * surfdisp_batch reads.  It replaces, for M models at once, the reference's per-model Python assembly:
 * B-spline basis x coefficients on the fine grid of every layer group (layers.py:4-45, 104-136), the per-class
 * Vp / density / Q rules (layers.py:139-295), the stacking of the groups (models.py:72-91), mid-point
 * averaging to layers and the h > 0.01 km filter (models.py:93-102, models.py:20).
 * A template describes the groups (top to bottom); every numeric entry is either fixed or refers to one of
 * the P free parameters of a model (column order = MCinv._brownians order, models.py:227-240).
 */
#define SURFDISP_MAX_GROUPS 8
#define SURFDISP_MAX_COEF 8

/* group kinds (Vs profile inside the group)                                   reference class            */
#define SURFDISP_G_WATER 0      /* Vs = 0                                        OceanWater  layers.py:191  */
#define SURFDISP_G_CONST 1      /* constant                                      Sediment / OceanSediment    */
#define SURFDISP_G_LINEAR 2     /* linear top..bottom                            Sediment / OceanCrust       */
#define SURFDISP_G_BSPLINE 3    /* B-spline, ncoef coefficients                  Crust / Mantle  :158, :239  */
#define SURFDISP_G_CASCADIA 4   /* (0.02 H^2 + 1.27 H + 0.029) / (H + 0.29)      OceanSedimentCascadia :289  */
#define SURFDISP_G_REFMANTLE 5  /* linear continuation below the model           ReferenceMantle :267        */
#define SURFDISP_G_HYBRID 6     /* thermal mantle: half-space-cooling temperature (ThermSeis.HSCM, ThermSeis.py:56-101)
                                   -> Vs by the mineral-physics relations of OceanSeisRitz (ThermSeis.py:103-176), B-spline
                                   perturbation (coefficients [0, v_1 .. v_ncoef]) below the depth where melting starts,
                                   joined by a not-a-knot cubic spline; Qs from OceanSeisRuan (ThermSeis.py:320-448)
                                                                                    OceanMantleHybrid layers.py:297-363 */
/* number of fine layers of a group */
#define SURFDISP_N_FIXED 0
#define SURFDISP_N_CRUST 1      /* 5/10/15/30/60 by thickness, layers.py:161-173 */
#define SURFDISP_N_OCRUST 2     /* min(max(round(H/2), 2), 10), layers.py:226 */
/* density rules */
#define SURFDISP_R_QUARTIC 0    /* quartic in Vs, layers.py:152 */
#define SURFDISP_R_OCEAN 1      /* 0.541 + 0.3601 Vp, layers.py:216 */
#define SURFDISP_R_MANTLE 2     /* 3.4268 + (Vs - 4.5) / 4.5, layers.py:262 */
#define SURFDISP_R_CONST 3
/* group classes (layers.py prop['Group']) used by the prior checks */
#define SURFDISP_C_WATER 0
#define SURFDISP_C_SEDIMENT 1
#define SURFDISP_C_CRUST 2
#define SURFDISP_C_MANTLE 3
#define SURFDISP_C_OTHER 4
/* prior checks of CascadiaPrism.isgood (models.py:294-360), bits of SurfdispStackTemplate.prior_mask and of the
 * per-model result */
#define SURFDISP_P_JUMP 1        /* Vs jump between groups must not be negative (models.py:305-307) */
#define SURFDISP_P_VSMAX 2       /* all Vs <= 4.9 km/s (models.py:312-313) */
#define SURFDISP_P_MONO 4        /* Vs strictly increasing inside sediment and crust (models.py:318-321) */
#define SURFDISP_P_BOTTOM 8      /* positive Vs gradient at the bottom of the mantle (models.py:355-356; for the ocean
                                    rules: at the bottom of the whole grid, models.py:597-598) */
/* CascadiaOcean.isgood (models.py:571-677), evaluated when prior_mask asks for any of the last five */
#define SURFDISP_P_SEDMIN 16     /* Vs in the sediment >= 0.2 km/s (models.py:581-583) */
#define SURFDISP_P_FIRSTPAIR 32  /* what is left of the jump rule there: `grp` is a Python list in that method, so only
                                    the first two grid points are compared (models.py:586-588) */
#define SURFDISP_P_OSCI 64       /* neighbouring local extrema of the mantle Vs differ by < 10 % of its mean (:603-611) */
#define SURFDISP_P_LOCALMAX 128  /* no local maximum in the mantle (:616-619) */
#define SURFDISP_P_SLOPE 256     /* no gradient below 1.5 x the gradient under the moho (:621-623) */
#define SURFDISP_P_CWT 512       /* neighbouring extrema of the Mexican-hat transform (width 30 km) of the detrended
                                    mantle profile differ by <= 0.3 (:626-635; scipy.signal.cwt / ricker of SciPy <= 1.14) */
/* the rule sets of the reference's model classes */
#define SURFDISP_PRIOR_PRISM (1 | 2 | 4 | 8)                        /* CascadiaPrism.isgood      models.py:294-360 */
#define SURFDISP_PRIOR_CONTINENT (1 | 2 | 4)                        /* CascadiaContinent.isgood  models.py:385-523 */
#define SURFDISP_PRIOR_OCEAN (16 | 32 | 8 | 64 | 128 | 256 | 512)   /* CascadiaOcean.isgood      models.py:571-677 */

typedef struct SurfdispStackGroup {
  int kind, nfine_rule, nfine, h_mode;      /* h_mode 0: parameter is the thickness H, 1: BottomDepth */
  int h_param, ncoef, rho_rule, gclass;     /* h_param: index of the free parameter or -1 (h_fixed);
                                               gclass: SURFDISP_C_* group of the reference (layer.prop['Group']) */
  int v_param[SURFDISP_MAX_COEF];           /* per Vs coefficient: free-parameter index or -1 (v_fixed) */
  double v_fixed[SURFDISP_MAX_COEF];
  double h_fixed, vp_a, vp_b, rho_const, qs, slope;   /* Vp = vp_a Vs + vp_b; slope: km/s per km (REFMANTLE) */
  /* SURFDISP_G_HYBRID only */
  int age_param, pad_;                      /* ThermAge: free-parameter index or -1 (age_fixed), Myr */
  double age_fixed, tp, period, q_age;      /* potential temperature (deg C, 1325); period of the Q model (Info.period, 1 s);
                                               age of the Q model (Info.lithoAge when Info.lithoAgeQ), < 0: the ThermAge */
} SurfdispStackGroup;

typedef struct SurfdispStackTemplate {
  int ngroups, nparams;
  int prior_mask, pad_;                      /* which SURFDISP_P_* rules the Monte-Carlo walk enforces (0: none, MCinv.isgood) */
  double topo;                               /* km, negative below sea level (models.py:76) */
  SurfdispStackGroup groups[SURFDISP_MAX_GROUPS];
} SurfdispStackTemplate;

/* The host-buffer pipeline fed with PARAMETER vectors instead of layers (Model1D.forward for a batch: models.py:93-121
 * = seisPropLayers + _calForward): params HOST float[M][nparams] (pinned for an asynchronous copy) are copied once,
 * the stacks are assembled on the device (surfdisp_build_stacks), then the stages of surfdisp_host_batch_pipelined
 * follow.  Host<->device traffic per model: 4 nparams bytes in, 8 K + 8 bytes out -- the layer arrays (20 Lmax bytes
 * per model) never cross PCIe.  device_buffer: surfdisp_params_pipelined_bytes() bytes. */
size_t surfdisp_params_pipelined_bytes(int n_models, int n_params, int n_layers_max, int n_periods);
int surfdisp_host_params_pipelined(const SurfdispOpts* opts, const SurfdispStackTemplate* tmpl, int kind, int n_models,
                                   int n_layers_max, const float* params, int n_periods, const float* periods,
                                   float* c_out, float* u_out, int* nfound, int* flags, void* device_buffer,
                                   size_t device_bytes, int n_chunks, void* compute_stream, void* copy_stream);

/* params: device float[M][nparams]; layers: device float[5][M][n_layers_max]; n_layers: device int[M]
 * (0 if a stack would need more than n_layers_max layers).  Asynchronous on `stream`. */
int surfdisp_build_stacks(const SurfdispStackTemplate* tmpl, int n_models, const float* params,
                          int n_layers_max, float* layers, int* n_layers, void* stream);

/* Prior checks only: priors device int[M] = SURFDISP_P_* bits violated (all rules are evaluated, whatever
 * prior_mask says). */
int surfdisp_check_priors(const SurfdispStackTemplate* tmpl, int n_models, const float* params, int* priors,
                          void* stream);

/* One Monte-Carlo proposal per chain, replacing MCinv.perturb / reset (models.py:190-219) with
 * BrownianVar.move / reset (brownian.py:17-27): every free parameter gets a Gaussian step, redrawn (up to 1000
 * times, then a uniform redraw) until it is inside (vmin, vmax); the whole proposal is redrawn (up to 1000 times,
 * then uniform resets, up to 10000) until the enabled priors hold.
 *   lo, hi, step   HOST float[P]  (BrownianVar vmin, vmax, step)
 *   cur            device float[M][P] current models; reset_mask device unsigned char[M] or NULL: chains that
 *                  start over from a uniform draw (chain restarts, point.py:47-57)
 *   prop           device float[M][P] proposals (out); status device int[M] or NULL: tries used, negative if no
 *                  admissible model was found (the reference raises there)
 *   seed, step_index  counter-based generator (Philox4x32-10): chain m at step s always draws the same numbers
 */
int surfdisp_mc_propose(const SurfdispStackTemplate* tmpl, int n_chains, const float* lo, const float* hi,
                        const float* step, const float* cur, const unsigned char* reset_mask, float* prop,
                        int* status, unsigned long long seed, unsigned int step_index, void* stream);

/* Metropolis rule of point.py:34-37 for every chain: accept if chi1 < chi0, else if u > 1 - exp(-(chi1-chi0)/2).
 * chi0 / cur are updated in place where the proposal is accepted; accepted device unsigned char[M] (out).
 * force_mask device unsigned char[M] or NULL: chains whose proposal is taken unconditionally (first sample of a
 * chain, point.py:57). */
int surfdisp_mc_accept(int n_chains, int n_params, const float* chi1, const float* prop, float* chi0, float* cur,
                       const unsigned char* force_mask, unsigned char* accepted, unsigned long long seed,
                       unsigned int step_index, void* stream);

/* ---- One whole Monte-Carlo step of an ensemble of chains, every stage on the device (Point.MCinv, point.py:40-76,
 * for M chains at once; the chains of several points -- the grid nodes of model3D.py:50-57 -- run side by side):
 *   proposal + model assembly (one kernel, warp per chain) -> surfdisp_batch stages (phase velocities) -> misfit +
 *   Metropolis rule + state update + chain-track row (one kernel) -> step counter + 1.
 * All launches go to `stream` and read the step index from a device counter, so a captured CUDA graph of one call
 * can be replayed for the following steps.
 *   step % chain_len == 0 starts a sub-chain (point.py:45-57): at step 0 the chains flagged in init_mask take their
 *   current model as it is (perturbed first if it violates the priors); every other start is a uniform redraw over the
 *   box; the first sample of a sub-chain is always kept. */
typedef struct SurfdispMcState {
  int n_chains, n_params, n_periods, n_layers_max;
  int kind;                 /* SURFDISP_KIND_* of the observed curve */
  int misfit_mode;          /* 0 Point.misfit (point.py:15-31), 1 PointCascadia.misfit (point.py:337-366) */
  int chain_len;            /* sub-chain length (chainL of point.py:32); 0: no restarts */
  int chains_per_point;     /* chain m belongs to point m / chains_per_point */
  int track_steps;          /* rows of the track ring buffer (0: no track) */
  int pad_;
  unsigned long long seed;
  /* device pointers */
  float* cur;               /* [M][P] current models (in/out) */
  float* prop;              /* [M][P] proposals of this step (out) */
  float* chi0;              /* [M] chi-square of the current models (in/out) */
  int* status;              /* [M] tries of the proposal, < 0: no admissible model found (the reference raises there;
                               such a chain keeps its state and the row is written with accepted = 0) */
  unsigned char* accepted;  /* [M] (out) */
  const unsigned char* init_mask;   /* [M] or NULL */
  float* misfit;            /* [M][3] (misfit, chiSqr, L) of the proposals (out) or NULL */
  float* track;             /* [track_steps][M][3 + P] rows [misfit, L, accepted, parameters] (models.py:243-245) or NULL */
  unsigned int* step;       /* [1] step counter (in/out) */
  const float* bounds;      /* [n_points][3][64]: vmin, vmax, step of every free parameter (brownian.py:3-16) */
  const float* obs;         /* [n_points][K] observed phase velocities */
  const float* isig;        /* [n_points][K] 1 / uncertainty */
  const unsigned char* use; /* [n_points][K] 0 = masked observation */
  /* scratch owned by the caller */
  float* layers; int* n_layers;          /* [5][M][n_layers_max], [M] */
  float* c_pred; int* nfound; int* flags;   /* [M][K], [M], [M] */
  float* c_cur;                             /* [M][K] curve of every chain's current model (zero-initialised by the caller;
                                               kept by the step: the neighbour curve of surfdisp_batch_hinted) or NULL */
  void* workspace; size_t workspace_bytes;  /* surfdisp_workspace_bytes(M, n_layers_max, K) */
} SurfdispMcState;

int surfdisp_mc_step(const SurfdispOpts* opts, const SurfdispStackTemplate* tmpl, const SurfdispMcState* st,
                     const float* periods /* HOST float[K] */, void* stream);

const char* surfdisp_version(void);
const char* surfdisp_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif
