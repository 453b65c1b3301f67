/*
 * spamm_b200 -- C ABI of the B200-native SPAMM ln-posterior hot path.
 *
 * The reference (jotaylor/SPAMM) is pure Python and has no FFI; this header is
 * the boundary a maintainer would bind with ctypes.  Each
 * entry point names the reference interface it replaces (paths relative to
 * the reference root).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success or a negative SPAMM_E* code (message via spamm_last_error()); no C++
 * exceptions cross the ABI.  `*_dev` pointers are CUDA device pointers owned by
 * the caller (torch tensors in the Python host layer); `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Numeric edge cases
 * are never errors: out-of-prior -> -inf, NaN likelihood -> 0.0, +-inf ->
 * +-DBL_MAX, exactly like spamm/Model.py:67-69,443.
 */
#ifndef SPAMM_B200_H
#define SPAMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPAMM_ABI_VERSION 4
#define SPAMM_MAX_COMPONENTS 8
#define SPAMM_MAX_TEMPLATES 16
#define SPAMM_SH95_NE 7
#define SPAMM_SH95_T 10
#define SPAMM_SH95_LINES 48

/* error codes */
#define SPAMM_OK 0
#define SPAMM_EINVAL (-1)     /* bad argument / unsupported configuration */
#define SPAMM_ECUDA (-2)      /* CUDA runtime error */
#define SPAMM_ENOMEM (-3)
#define SPAMM_ERANGE (-4)     /* a walker needed a value outside the plan's tables, or a peer rank's
                                 results never arrived: the numbers of that call are not the model's */

/* Model.components entries, spamm/run_spamm.py:100-122 */
enum SpammComponentKind {
    SPAMM_COMP_NUCLEAR = 0,  /* NuclearContinuumComponent */
    SPAMM_COMP_FE = 1,       /* FeComponent               */
    SPAMM_COMP_HOST = 2,     /* HostGalaxyComponent       */
    SPAMM_COMP_BALMER = 3,   /* BalmerCombined            */
    SPAMM_COMP_EXTINCTION = 4 /* Extinction (spamm/components/ReddeningLaw.py:126-216); must be last */
};

enum SpammLineType { SPAMM_LINE_LORENTZIAN = 0, SPAMM_LINE_GAUSSIAN = 1 };

/* How the per-walker Gaussian broadening of Fe/host templates is evaluated. */
enum SpammTemplateMode {
    SPAMM_TEMPLATES_AUTO = 0,
    SPAMM_TEMPLATES_BANK = 1,   /* exact pre-broadened bank: sigma_norm is an integer
                                   (np.ceil, FeComponent.py:290), built once by the
                                   broadening kernel for every sigma the prior admits */
    SPAMM_TEMPLATES_DIRECT = 2  /* broadening kernel run per walker per call */
};

/* One template family on its uniform ln-lambda grids, i.e. the state that
 * FeComponent.initialize (FeComponent.py:156-191) / HostGalaxyComponent.initialize
 * (HostGalaxyComponent.py:154-188) leave in self.log_fe / self.log_host. */
typedef struct SpammTemplateSet {
    int32_t n_templates;                       /* <= SPAMM_MAX_TEMPLATES */
    int32_t n_points[SPAMM_MAX_TEMPLATES];     /* N_t */
    const double* log_wl;                      /* concatenated ln-lambda grids, sum(N_t) */
    const double* log_flux;                    /* concatenated fluxes on those grids    */
    double norm_wl[SPAMM_MAX_TEMPLATES];       /* np.median(template wavelengths), Fe:325 / HG:330 */
    double template_width;                     /* fe_template_width / hg_template_stellar_disp */
    double sigma_denom;                        /* c_kms * 2 sqrt(2 ln 2), Fe:286 / HG:293 */
    double sqrt_2pi;                           /* np.sqrt(2*math.pi), Fe:293 */
    int32_t kernel_size_sigma;                 /* fe_kernel_size_sigma / hg_kernel_size_sigma */
    int32_t clamp_edges;                       /* 0: np.interp(left=0,right=0) (Fe:315-319); 1: edge clamp (HG:323-325) */
    /* rebin_spec: True (parameters.yaml; FeComponent.py:320-323 / HostGalaxyComponent.py:326-329): the
     * convolved row goes onto the data grid by the flux-conserving rebin instead of np.interp.  The rebin is
     * linear, so the host hands over its matrix per template in CSR form: rebin_indptr is
     * [n_templates][P + 1] with offsets into rebin_index / rebin_weight (template-local column indices).
     * All NULL: np.interp (the default, rebin_spec: False). */
    const int32_t* rebin_indptr;
    const int32_t* rebin_index;
    const double* rebin_weight;
} SpammTemplateSet;

/* Frozen snapshot of a Model after `model.data_spectrum = spectrum` and after
 * the string prior sentinels were resolved.  All
 * arrays are HOST pointers, copied during spamm_plan_create. */
typedef struct SpammPlanDesc {
    int32_t abi_version;                       /* SPAMM_ABI_VERSION */
    int32_t n_pix;                             /* P */
    const double* wl;                          /* [P] data_spectrum.spectral_axis, strictly ascending */
    const double* ln_wl;                       /* [P] np.log(wl) as the host computed it */
    const double* flux;                        /* [P] data_spectrum.flux (NULL for flux-only plans) */
    const double* flux_error;                  /* [P] data_spectrum.flux_error                       */
    double sum_ln_2pi_err2;                    /* np.sum(np.log(2 pi err^2)), Model.py:440 */

    int32_t n_components;                      /* Model.components, in order (Model.py:353-364) */
    int32_t component_kind[SPAMM_MAX_COMPONENTS];
    int32_t n_dim;                             /* D = total_parameter_count (Model.py:295-308) */
    const double* prior_lo;                    /* [D] strict flat-box priors (NC:148-170, Fe:241-248, */
    const double* prior_hi;                    /* [D]  HG:232-248, BC:161-189); may be NULL => no prior */

    /* NuclearContinuumComponent (NuclearContinuumComponent.py:176-201) */
    int32_t nc_broken_pl;                      /* 0: [norm_PL, slope1]; 1: [wave_break, norm_PL, slope1, slope2] */
    double nc_norm_wl;                         /* Spectrum.norm_wavelength = np.median(wl), Spectrum.py:84-87 */

    SpammTemplateSet fe;                       /* used when a SPAMM_COMP_FE is present   */
    SpammTemplateSet host;                     /* used when a SPAMM_COMP_HOST is present */
    int32_t template_mode;                     /* SpammTemplateMode */

    /* BalmerCombined (BalmerContinuumCombined.py) */
    int32_t balmer_bc;                         /* BalmerContinuum flag   (:83)  */
    int32_t balmer_bpc;                        /* BalmerPseudocContinuum (:84)  */
    int32_t balmer_line_type;                  /* SpammLineType, bc_line_type   */
    int32_t balmer_line_min;                   /* int(bc_lines_min) = 3         */
    int32_t balmer_n_lines;                    /* len(np.arange(bc_lines_min, bc_lines_max)) = 397 */
    int32_t balmer_exact_line_sum;             /* 0: far-field cluster series for distant line groups
                                                  (<= 1e-15 relative, default); 1: every line summed
                                                  directly like makelines (BC:315-333) */
    const double* balmer_line_center;          /* [n_lines] balmerseries(), :247-259 */
    const double* balmer_line_exparg;          /* [n_lines] E0*(1/n^2 - 1/(n-1)^2), :296 */
    double sh95_ne[SPAMM_SH95_NE];             /* knots of the 48 pickled interp2d objects, :288-290 */
    double sh95_T[SPAMM_SH95_T];
    const double* sh95_z;                      /* [48][7][10]; list index k <-> coef_use[k] */
    const double* bc_ln_wl_new;                /* [P] log_conv's uniform ln-lambda grid, :219-221 */
    const double* bc_exp_ln_wl_new;            /* [P] np.exp(ln_wavenew), :239 */
    /* physical constants exactly as the host holds them (astropy 3.x, CODATA-2014) */
    double c_kms, c_cgs, c_aa, h_cgs, kb_cgs, balmer_edge;

    int32_t max_walkers;                       /* workspace is sized for this many walkers per call */

    /* Extinction (ReddeningLaw.py): k(lambda) [P] of the selected law as the host evaluated it
     * (Calzetti_ext :7-20, LMC :22-46, MW :48-72, SMC :74-98, AGN :100-123; zeros = no law);
     * the kernel multiplies the flux summed so far by pow(10, -0.4 * E(B-V) * k), Model.py:384-395.
     * One parameter, E(B-V).  NULL unless a SPAMM_COMP_EXTINCTION is present. */
    const double* ext_k;

    /* rebin_spec: True for BalmerCombined.log_conv (BalmerContinuumCombined.py:227,241): the two rebins
     * around the (identity) convolution composed into one sparse P x P operator on the unnormalised
     * continuum, CSR: bc_rebin_indptr [P + 1], bc_rebin_index, bc_rebin_weight.  All NULL: the two np.interp
     * resamplings (the default), which the library collapses into a 4-tap stencil itself. */
    const int32_t* bc_rebin_indptr;
    const int32_t* bc_rebin_index;
    const double* bc_rebin_weight;
} SpammPlanDesc;

typedef struct SpammPlan SpammPlan;

/* Build the device-resident plan: uploads the snapshot, runs the template
 * broadening kernel to fill the exact sigma bank, precomputes the log_conv
 * resampling operator.  Replaces the per-call setup the reference repeats in
 * every flux() (FeComponent.py:285-293, BalmerContinuumCombined.py:218-221,288). */
int spamm_plan_create(const SpammPlanDesc* desc, SpammPlan** out_plan);
void spamm_plan_destroy(SpammPlan* plan);

/* ln_posterior(new_params, model) for a whole walker ensemble in one call:
 * spamm/Model.py:42-79 (prior :449-470, model_flux :328-366, likelihood
 * :418-445).  params_dev is [n_walkers][D] row-major, lnp_dev is [n_walkers].
 * Asynchronous on `stream`. */
int spamm_lnpost_batch(SpammPlan* plan, const double* params_dev, int32_t n_walkers,
                       double* lnp_dev, void* stream);

/* Same call with HOST buffers: pinned staging, H2D, kernel, D2H, synchronise.
 * This is what an emcee `pool.map` adapter binds (Model.py:283-289).
 * Returns SPAMM_ERANGE (and clears the condition) if any walker set the device status word. */
int spamm_lnpost_batch_host(SpammPlan* plan, const double* params_host, int32_t n_walkers,
                            double* lnp_host);

/* Model.model_flux (Model.py:328-366) restricted to the components whose bit is
 * set in component_mask (bit i = Model.components[i]); flux_dev is
 * [n_walkers][P].  With one bit set this is Component.flux(spectrum, parameters)
 * (ComponentBase.py:82-84) batched over walkers; params_dev always holds the
 * full [n_walkers][D] vectors.  No prior test is applied. */
int spamm_model_flux_batch(SpammPlan* plan, uint32_t component_mask, const double* params_dev,
                           int32_t n_walkers, double* flux_dev, void* stream);

/* Model.likelihood (Model.py:418-445) for model fluxes already on the device:
 * ln_l = nan_to_num(-0.5 * sum(((flux - model)/err)^2 + ln(2 pi err^2))).
 * model_flux_dev is [n_walkers][P], lnl_dev is [n_walkers]. */
int spamm_likelihood_batch(SpammPlan* plan, const double* model_flux_dev, int32_t n_walkers,
                           double* lnl_dev, void* stream);

/* Query helpers */
int spamm_plan_n_dim(const SpammPlan* plan);
int spamm_plan_n_pix(const SpammPlan* plan);
int spamm_plan_template_mode(const SpammPlan* plan);          /* resolved SpammTemplateMode */
int spamm_plan_bank_rows(const SpammPlan* plan);              /* rows in the sigma bank */
/* 1 if ln-posterior calls on this plan run the compile-time specialised instantiation of the fused
 * kernel (the headline model: NC + Fe x3 + host x7 + BalmerCombined, Lorentzian lines, bank rows, a
 * prior box bounding every exp argument, even P), 0 if the generic one. */
int spamm_plan_specialised(const SpammPlan* plan);
/* Device status word (0 = ok); synchronises `stream`.  Non-zero means some walker
 * needed a table entry outside the plan (SPAMM_ERANGE conditions). */
int spamm_plan_status(SpammPlan* plan, void* stream);
/* Same check as an error code: synchronises `stream`, returns SPAMM_ERANGE with the conditions spelled
 * out in spamm_last_error() and clears the word, or SPAMM_OK.  The synchronous entry points
 * (spamm_lnpost_batch_host, spamm_lnpost_batch_sharded_host) make this check themselves; the
 * asynchronous ones queue a copy of the word behind their kernels and the NEXT call on the plan
 * returns the error -- call spamm_plan_check where you synchronise (end of a chain) to be sure. */
int spamm_plan_check(SpammPlan* plan, void* stream);
/* number of kernels the library has launched on this plan since creation */
int64_t spamm_plan_launch_count(const SpammPlan* plan);

/* Small ensembles (emcee's default is 30 walkers, run_spamm.py:26): with fewer walkers than
 * resident CTA slots the fused kernel runs one walker per thread-block CLUSTER and deals the
 * walker's pixel tasks over the cluster's CTAs; the chi^2 partials are gathered through
 * distributed shared memory in the same fixed order, so results do not depend on the split.
 * Ensembles of one to a few waves of CTAs (a half-ensemble shard on one of 8 GPUs) can instead run each
 * walker as two independent CTAs -- the pixels redward of the Balmer edge with the line table, the pixels
 * blueward with the Balmer continuum stage -- with dynamically dealt tasks; same bits again.
 * mode 0: off; 1: automatic (default); 2, 4, 8: forced cluster size; 3: forced two-part split. */
int spamm_plan_set_walker_split(SpammPlan* plan, int32_t mode);

/* emcee 2.2.1 stretch move (EnsembleSampler._propose_stretch; call site
 * Model.py:283-289), one half-ensemble at a time, counter-based Philox RNG.
 *   propose: q[i] = c[r_i] - z_i (c[r_i] - s[i]),  z_i = ((a-1) u + 1)^2 / a
 *   accept : (D-1) ln z + lnp(q) - lnp(s) > ln u'
 * walkers_dev [K][D], lnp_dev [K]; half = 0 moves walkers [0,K/2) using
 * [K/2,K) as the complementary set, half = 1 the reverse.  q_dev [K/2][D],
 * z_dev [K/2].  chain_dev/lnp_chain_dev may be NULL; otherwise the state is
 * recorded at [walker][iter][*] (layout of emcee's sampler.chain). */
int spamm_stretch_propose(const double* walkers_dev, int32_t n_walkers, int32_t n_dim, int32_t half,
                          double a, uint64_t seed, uint64_t iteration, double* q_dev, double* z_dev,
                          void* stream);
int spamm_stretch_accept(double* walkers_dev, double* lnp_dev, int32_t n_walkers, int32_t n_dim,
                         int32_t half, const double* q_dev, const double* z_dev,
                         const double* lnp_q_dev, uint64_t seed, uint64_t iteration,
                         int64_t* naccepted_dev, void* stream);
int spamm_chain_record(const double* walkers_dev, const double* lnp_dev, int32_t n_walkers,
                       int32_t n_dim, int64_t iter, int64_t n_iter, double* chain_dev,
                       double* lnp_chain_dev, void* stream);

/* n_iter whole iterations of the stretch move on one device -- for each iteration both half-steps
 * (propose, spamm_lnpost_batch on the proposals, accept) and, if chain_dev/lnp_chain_dev are not
 * NULL, the chain record at [walker][iter0 + i][*] of a chain of length chain_len -- queued on
 * `stream` in one call.  This is the loop of emcee's EnsembleSampler.sample() (call site
 * Model.py:283-289) without a round trip to the host language per half-step.  q_dev [K/2][D],
 * z_dev [K/2], lnp_q_dev [K/2] are caller-owned scratch. */
int spamm_stretch_run(SpammPlan* plan, double* walkers_dev, double* lnp_dev, int32_t n_walkers, double a,
                      uint64_t seed, uint64_t iter0, int32_t n_iter, double* q_dev, double* z_dev,
                      double* lnp_q_dev, int64_t* naccepted_dev, double* chain_dev, double* lnp_chain_dev,
                      int64_t chain_len, void* stream);

/* ---- multi-GPU: walker shards, results exchanged by the compute kernel itself -------------------
 * Replaces the task farm of emcee's pool (spamm/Model.py:264-289: pickled (Model, params) tasks out,
 * floats back).  Every rank holds the replicated ensemble and the same counter-based RNG; a sharded call
 * evaluates walkers [lo, hi) = spamm_shard_bounds(n, n_ranks, rank) on this rank: the walker kernel leaves
 * each ln-posterior in this rank's exchange region and, when its last walker retires, raises this rank's
 * flag in EVERY rank's region (release stores over NVLink).  Consumers (the gather kernel, the sampler's
 * accept kernel) wait on the flags on the device and pull the other ranks' entries straight out of their
 * regions (peer loads over NVLink); there is no host-side barrier and no collective on the data path.
 * All ranks must make the same sequence of sharded calls (the epoch counter is implicit). */
typedef struct SpammExchange SpammExchange;
#define SPAMM_IPC_HANDLE_BYTES 64
void spamm_shard_bounds(int64_t n, int32_t n_ranks, int32_t rank, int64_t* lo, int64_t* hi);
int64_t spamm_exchange_region_bytes(int64_t max_walkers);
/* Allocate this rank's region on the current device (flag block + two result buffers of max_walkers
 * doubles).  ipc_handle_out (SPAMM_IPC_HANDLE_BYTES, may be NULL) receives the CUDA IPC handle the other
 * processes of the box need; n_ranks == 1 is connected immediately. */
int spamm_exchange_create(int32_t n_ranks, int32_t rank, int64_t max_walkers, SpammExchange** out,
                          void* ipc_handle_out);
/* all_handles: [n_ranks][SPAMM_IPC_HANDLE_BYTES] gathered from every rank by any transport
 * (torch.distributed.all_gather, MPI, a file).  Peer access is enabled as the handles are opened.
 * The caller must barrier after connecting and before destroying. */
int spamm_exchange_connect(SpammExchange* ex, const void* all_handles);
/* Same-process variant (several ranks as threads/streams of one process, or regions mapped by the
 * caller, e.g. torch symmetric memory): region_ptrs[r] = rank r's region as addressable from here. */
int spamm_exchange_connect_ptrs(SpammExchange* ex, const uint64_t* region_ptrs);
uint64_t spamm_exchange_region_ptr(const SpammExchange* ex);
uint64_t spamm_exchange_epoch(const SpammExchange* ex);
void spamm_exchange_destroy(SpammExchange* ex);

/* spamm_lnpost_batch over a sharded ensemble: params_dev is the FULL replicated [n_total][D] array
 * (only this rank's rows are read).  Queues on `stream`: the walker kernel on this rank's slice (results
 * into this rank's region, flag to every rank), then the gather kernel: wait for every rank's flag, pull
 * the other ranks' slices, and -- if lnp_dev != NULL -- write the gathered [n_total] vector to lnp_dev too.
 * *lnp_region_out (may be NULL) is the vector inside this rank's region; it stays valid until the
 * second-next sharded call. */
int spamm_lnpost_batch_sharded(SpammPlan* plan, SpammExchange* ex, const double* params_dev, int32_t n_total,
                               double* lnp_dev, const double** lnp_region_out, void* stream);
/* Host-buffer form: H2D of this rank's parameter rows only, kernel + exchange, D2H of the whole gathered
 * vector, synchronise, status check.  params_host is the full [n_total][D] array on every rank. */
int spamm_lnpost_batch_sharded_host(SpammPlan* plan, SpammExchange* ex, const double* params_host,
                                    int32_t n_total, double* lnp_host);
/* spamm_stretch_run with every half-step's proposals sharded over the ranks: propose (replicated), walker
 * kernel on this rank's slice (+ flag), accept after the flag wait inside the accept kernel, which pulls
 * each proposal's ln-posterior from the rank that evaluated it.
 * Chains are bit-identical to the single-device run with the same seed. */
int spamm_stretch_run_sharded(SpammPlan* plan, SpammExchange* ex, double* walkers_dev, double* lnp_dev,
                              int32_t n_walkers, double a, uint64_t seed, uint64_t iter0, int32_t n_iter,
                              double* q_dev, double* z_dev, int64_t* naccepted_dev, double* chain_dev,
                              double* lnp_chain_dev, int64_t chain_len, void* stream);

/* Timing probe of the template broadening kernels (FeComponent.py:285-319 / HostGalaxyComponent.py:292-325
 * replaced by k_conv_rows + k_interp_rows + k_norm_rows) on their own: n_rows rows of family 0 (Fe) / 1 (host),
 * every row with integer kernel width sigma_norm; mean durations over `repeats` launches (CUDA events) and the
 * FMA count of one convolution launch.  Roofline evidence for the per-walker (DIRECT) broadening path. */
int spamm_broadening_probe(SpammPlan* plan, int32_t family, int32_t sigma_norm, int32_t n_rows, int32_t repeats,
                           double* conv_ms_out, double* interp_ms_out, double* conv_fma_out);

/* FP64 FMA throughput probe (roofline denominator: MEASURED_PEAKS.json has no
 * FP64 entry).  Runs `iters` dependent-chain DFMA iterations on every SM and
 * returns the achieved TFLOP/s (2 flop per FMA) measured with CUDA events. */
int spamm_fp64_peak_probe(int32_t iters, int32_t repeats, double* tflops_out);

const char* spamm_last_error(void);
int spamm_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SPAMM_B200_H */
