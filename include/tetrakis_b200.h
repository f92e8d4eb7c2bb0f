/*
 * tetrakis_b200.h -- C ABI of the B200-native wave-propagation hot path of tetrakis-sim.
 *
 * The reference (pure Python) has no FFI layer; its boundary for this path is two functions,
 *     tetrakis_sim.physics.run_wave_sim   (reference tetrakis_sim/physics.py:41-131)
 *     tetrakis_sim.physics.run_fft        (reference tetrakis_sim/physics.py:140-165)
 * This header is what a ctypes / cffi / CPython-extension binding inside those two functions
 * would call (INTEGRATION.md shows the stub).  Plain pointers and sizes only; every function
 * returns an int status (TK_OK == 0) and leaves a message for tk_last_error() on failure.
 *
 * Ownership: caller owns every host buffer passed in or out; the library owns all device
 * memory, released by tk_plan_destroy().  A plan is bound to one device and one stream and is
 * not thread-safe; distinct plans are independent.  There is NO CPU fallback: every entry point
 * that computes fails with TK_ERR_NO_DEVICE when no CUDA device is present.
 *
 * Node numbering
 *   graph plans   : 0..n-1 in list(G.nodes) order (physics.py:56).
 *   lattice plans : linear index ((z*size + r)*size + c)*4 + q, q = 0..3 for "ABCD"; the
 *                   reference's (r,c[,z],q) tuples (lattice.py:30,48-54) map onto it.  A plan that
 *                   owns the slab [z_begin, z_end) uses LOCAL indices with z relative to z_begin.
 *   row-band plans: one defect-free sheet sharded over GPUs (tk_sheet_*): ((r - row_begin)*size + c)*4 + q.
 *
 * Families of entry points: plans (tk_graph_plan_create, tk_lattice_plan_create[_ex], tk_sheet_band_plan_create),
 * state and probes (tk_plan_set_state*, tk_plan_get_state, tk_plan_set_probes, tk_plan_read_probes), stepping
 * (tk_plan_run; tk_lattice_step_* + halo/peer/flag calls for 3-D slabs across GPUs; tk_sheet_pass_* for row bands),
 * spectra (tk_dft_abs), ensembles (tk_ensemble_run), introspection (tk_plan_*_info / _count, tk_lattice_tiling).
 */
#ifndef TETRAKIS_B200_H
#define TETRAKIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TK_ABI_VERSION 1

#define TK_OK 0
#define TK_ERR_INVALID 1   /* bad argument */
#define TK_ERR_CUDA 2      /* CUDA runtime error (message has the cudaError string) */
#define TK_ERR_NOMEM 3     /* host or device allocation failed */
#define TK_ERR_NO_DEVICE 4 /* no usable CUDA device: there is no CPU path */

typedef struct tk_plan tk_plan;

int tk_abi_version(void);
/* Thread-local, NUL-terminated description of the last failure on this thread. */
const char *tk_last_error(void);
int tk_device_count(int *count);
int tk_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major,
                   int *cc_minor, int64_t *total_mem_bytes);
/* Sustained FP64 instruction rate and HBM bandwidth the pass-length heuristics of tk_plan_run assume for this device,
 * derived from its attributes (SM count, clocks, bus width) and the efficiencies measured on a B200. */
int tk_device_rates(int device, double *fp64_per_s, double *hbm_bytes_per_s);
/* The library keeps freed device blocks of up to 256 MB (at most 2 GB per device) for reuse by later plans, because
 * driver allocation calls cost more than a small simulation; this returns them to the driver.  TK_DEV_CACHE=0
 * disables the cache. */
int tk_device_trim(int device);

/* ------------------------------------------------------------------------------------------
 * Graph plan: any networkx graph.  Replaces the per-call precompute of physics.py:56-64,93-102
 * (node list, degrees, neighbour lists, inv_mass, potential).  Neighbours are summed in the
 * order given (adjacency order), which makes the device result bit-identical to the reference.
 *   rowptr[n+1], colidx[rowptr[n]] : neighbour lists (list(G[n]), physics.py:94)
 *   degree[n]                      : G.degree (physics.py:60-62; self-loops count twice)
 *   inv_mass[n], potential[n]      : physics.py:97-102 (mass <= 0 -> 1 is the caller's job)
 * ------------------------------------------------------------------------------------------ */
int tk_graph_plan_create(int device, int64_t n_nodes, const int64_t *rowptr,
                         const int32_t *colidx, const double *degree, const double *inv_mass,
                         const double *potential, tk_plan **out);

/* ------------------------------------------------------------------------------------------
 * Lattice plan: a tetrakis sheet (lattice.py:15-76) described by its shape plus the sparse set
 * of nodes / edges that a defect (defects.py) changed.  Nothing O(N) crosses the ABI.
 *   special_* : nodes that are removed, edge-pruned or carry a mass/potential tag
 *       special_index  GLOBAL linear index (z over all `layers`)
 *       special_flags  bit0 alive (0 = removed, defects.py:133), bit1 part (0 = all incident
 *                      edges pruned, defects.py:227-232)
 *       special_inv_mass / special_potential  physics.py:97-102 values for that node
 *   corr_*    : directed edge corrections relative to the clique model; node a gains
 *               sign*u[b] in its neighbour sum and sign in its degree (wedge: two entries with
 *               sign -1 for the A-B edge, defects.py:78-81,97-103).  GLOBAL indices.
 *   [z_begin, z_end) : the slab of layers this plan owns (0, layers for a single GPU).  One halo
 *               plane is kept on each side; see tk_lattice_halo_ptrs.
 * ------------------------------------------------------------------------------------------ */
typedef struct tk_lattice_desc {
    int32_t size;   /* cells per side */
    int32_t layers; /* total layers; 1 for a 2-D sheet */
    int32_t z_begin;
    int32_t z_end;
    int64_t n_special;
    const int64_t *special_index;
    const uint8_t *special_flags;
    const double *special_inv_mass;
    const double *special_potential;
    int64_t n_corr;
    const int64_t *corr_a;
    const int64_t *corr_b;
    const double *corr_sign;
} tk_lattice_desc;

int tk_lattice_plan_create(int device, const tk_lattice_desc *desc, tk_plan **out);

/* Same with option flags.  TK_PLAN_F32: optional fp32 mode -- the state is stored and updated in
 * fp32 (12 B per node-update instead of 24), row/column sums and the Laplacian stay fp64.  The ABI
 * keeps double for every host-side array; tolerance vs fp64: DESIGN.md section 4. */
#define TK_PLAN_F32 0x1u
int tk_lattice_plan_create_ex(int device, const tk_lattice_desc *desc, uint32_t flags, tk_plan **out);

/* ------------------------------------------------------------------------------------------
 * Common plan API
 * ------------------------------------------------------------------------------------------ */
/* Run on an existing cudaStream_t (e.g. torch's current stream) instead of the plan's own. */
int tk_plan_set_stream(tk_plan *plan, void *cuda_stream);
/* Local node count (lattice plans: owned planes only, removed nodes included). */
int tk_plan_num_nodes(const tk_plan *plan, int64_t *n);
/* max G.degree over existing nodes (physics.py:64) -- input of the CFL clamp (physics.py:74-91). */
int tk_plan_max_degree(tk_plan *plan, int64_t *max_degree);
/* u <- 0, u_prev <- 0 (physics.py:58), then u[index[i]] = value[i] (physics.py:30-38). */
int tk_plan_set_state_sparse(tk_plan *plan, int64_t n, const int64_t *index, const double *value);
/* Dense host -> device; u_prev may be NULL (zeros).  Values at removed nodes are forced to 0. */
int tk_plan_set_state(tk_plan *plan, const double *u, const double *u_prev);
/* Dense device -> host; either pointer may be NULL. */
int tk_plan_get_state(tk_plan *plan, double *u, double *u_prev);

typedef struct tk_run_args {
    int64_t steps;
    double coeff;   /* (c * effective_dt)^2, physics.py:105 */
    double damping; /* physics.py:125 */
    int64_t n_probes;
    const int64_t *probe_index; /* local node indices */
    double *probe_out;          /* host, (steps+1) x n_probes, row 0 = state before the first step */
    double *history_out;        /* host or NULL, (steps+1) x num_nodes (physics.py:104,129) */
    double *energy_out;         /* host or NULL, `steps` values: entry s is the staggered discrete energy
                                   between states s and s+1,
                                     1/2 sum_i [ m_i (u_i^{s+1}-u_i^s)^2 / coeff + u_i^{s+1} ((V u^s)_i - (L u^s)_i) ]
                                   (x c^2 for physical units); exactly conserved when damping = 0.
                                   A new diagnostic: the reference computes no energy. */
    double *elapsed_ms;         /* host or NULL: device time of the stepping loop (CUDA events) */
    double *step_kernel_ms;     /* host or NULL: summed CUDA-event time of the step kernels alone */
    int64_t history_stride;     /* 0 or 1: history_out holds every state; k > 1: snapshots of states
                                   0, k, 2k, ... (floor(steps/k)+1 rows) for runs whose full history
                                   would not fit anywhere */
    uint64_t flags;             /* TK_RUN_* bits */
} tk_run_args;

/* tk_run_args.flags: one step per launch (24 B per node update) even where K-step passes would apply --
 * what a slab of a multi-GPU run does, so single-GPU denominators of weak-scaling series use it. */
#define TK_RUN_NO_FUSE 0x1u

/* physics.py:106-129: `steps` leapfrog updates, synchronous.  The state stays on the device, so
 * consecutive calls continue the same simulation. */
int tk_plan_run(tk_plan *plan, const tk_run_args *args);

/* Probe recording for phase-driven runs (tk_lattice_step_*): the next `rows` finalise calls (one in
 * tk_lattice_step_begin, one per tk_lattice_step_end) each record the probes' values.  tk_plan_run
 * manages its own probes and overrides this. */
int tk_plan_set_probes(tk_plan *plan, int64_t n_probes, const int64_t *probe_index, int64_t rows);
/* Copy the rows recorded so far (rows_recorded x n_probes doubles) to the host; synchronises. */
int tk_plan_read_probes(tk_plan *plan, double *out, int64_t *rows_recorded);

/* Kernel launches issued by this plan since creation (bench.py's gpu_launches). */
int tk_plan_launch_count(const tk_plan *plan, int64_t *launches);

/* Defect-free 2-D sheets (lattice.py:30-73 with no defects.py call): the row/column clique sums then obey
 * closed recurrences, so tk_plan_run advances K steps of physics.py:106-129 per pass over the state
 * (32/K bytes per node update instead of 24), unless a full history or the energy series is requested.
 * Reports how the LAST tk_plan_run went: steps per pass (0 = step by step), how many of its steps ran in
 * fused passes (steps mod K >= 2 runs as one shorter pass, a single left-over step as an ordinary step),
 * rows per CTA and resident CTAs per SM of the pass kernel.
 * Environment: TK_LAT_FUSE = 0 (off) or the steps per pass, >= 2.  Default: the largest K at which the pass is
 * still bound by moving the state (or by its launches) rather than by the FP64 pipe of a B200 -- 48 without
 * damping and 24 with on sheets that do not fit L2, up to 1024 on small ones. */
int tk_plan_fused_info(tk_plan *plan, int32_t *steps_per_pass, int64_t *fused_steps, int32_t *rows_per_cta,
                       int32_t *ctas_per_sm);

/* Lattices WITH defects (defects.py:55-107, 110-135, 194-247) and 3-D lattices (lattice.py:68-72): tk_plan_run
 * advances them in K-step passes too, with the "cross" scheme -- the rows and columns that hold a special cell in
 * any layer (plus those of the probes) are stepped one step at a time together with every row / column sum, whose
 * bulk parts obey closed recurrences; the rest of the lattice (every (r, c) column of cells is then an independent
 * chain along z) takes ONE pass per K steps: 32 B per bulk node + 24 B x K per cross node instead of 24 B x K per
 * node.  Used when the plan owns the whole lattice, the cross covers < 40 % of it, the lattice has >= 2^21 nodes,
 * and neither a full history nor the energy series is requested; tk_plan_fused_info reports K and the fused steps.
 * This call adds how large the cross was (rows, columns, fraction of the lattice) and how many passes ran since the
 * bulk parts were last summed from the data (3-D: every TK_FX_RESYNC = 16 passes; 2-D: every pass).
 * Environment: TK_LAT_FUSE (0 off; 3-D passes come in 8 or 16 steps), TK_FX_MAXFRAC (percent), TK_FX_MIN_NODES. */
int tk_plan_cross_info(tk_plan *plan, int32_t *steps_per_pass, int32_t *cross_rows, int32_t *cross_cols,
                       double *cross_fraction, int64_t *passes_since_anchor);

/* ------------------------------------------------------------------------------------------
 * ONE defect-free sheet sharded over GPUs by rows (one process per GPU).  Rank g holds rows
 * [row_begin, row_end) of the size x size sheet (node index = ((r - row_begin)*size + c)*4 + q).  Row sums are
 * local; the column sums and quadrant totals of a band are partial, so between passes they are summed over the
 * ranks: tk_sheet_pass_local leaves this band's contribution in `exchange` (device memory of
 * tk_sheet_exchange_doubles() doubles, e.g. a torch tensor), the caller all-reduces it (NCCL, SUM) on the
 * plan's stream, tk_sheet_pass_commit installs the global sums.  One collective per K steps:
 *
 *     tk_sheet_pass_local(plan, 0, coeff, damping, ex); allreduce(ex); tk_sheet_pass_commit(plan, ex);   // sums of u_0, u_{-1}
 *     repeat: tk_sheet_pass_local(plan, K, coeff, damping, ex); allreduce(ex); tk_sheet_pass_commit(plan, ex);
 *
 * State, probes (tk_plan_set_probes / tk_plan_read_probes; row 0 is recorded by the k = 0 call) and the stream
 * are handled with the generic tk_plan_* calls; tk_plan_run is not available on these plans.
 * Replaces: physics.py:106-129 on lattice.py:30-73 sheets too large (or too slow) for one GPU. */
int tk_sheet_band_plan_create(int device, int32_t size, int32_t row_begin, int32_t row_end, tk_plan **out);
int tk_sheet_exchange_doubles(tk_plan *plan, int64_t *count);
int tk_sheet_pass_local(tk_plan *plan, int32_t k, double coeff, double damping, void *exchange);
int tk_sheet_pass_commit(tk_plan *plan, const void *exchange);
/* Summed CUDA-event time of the pass kernels (lat_fused2d_kernel alone) launched since the previous call, at most
 * the 64 most recent, and how many they were; synchronises the plan's stream.  For bench.py's roofline. */
int tk_sheet_kernel_ms(tk_plan *plan, double *total_ms, int64_t *passes);
int tk_plan_destroy(tk_plan *plan);

/* ------------------------------------------------------------------------------------------
 * Slab stepping for multi-GPU runs (one process per GPU; the halo exchange itself is done by
 * the caller with NCCL / peer copies between the two phase calls).  All asynchronous on the
 * plan's stream.
 * ------------------------------------------------------------------------------------------ */
/* Device addresses of the four planes involved in a halo exchange: send_lo/send_hi = first/last
 * owned plane, recv_lo/recv_hi = halo below/above; plane_bytes = size*size*4*8.
 * next_state = 0: planes of the current state u_t; 1: planes of the buffer the running step
 * writes u_{t+1} into (so the exchange can start as soon as the boundary layers are updated and
 * overlap the interior layers).  The two buffers alternate at every tk_lattice_step_end. */
int tk_lattice_halo_ptrs(tk_plan *plan, int32_t next_state, void **send_lo, void **send_hi,
                         void **recv_lo, void **recv_hi, int64_t *plane_bytes);
/* Prepare a run: zero the probe cursor, compute the row/column sums of the current state. */
int tk_lattice_step_begin(tk_plan *plan, double coeff, double damping);
/* Update local layers [zl_begin, zl_end) (reads the current halos). */
int tk_lattice_step_layers(tk_plan *plan, int32_t zl_begin, int32_t zl_end);
/* Same, launched on another stream (e.g. a high-priority side stream for the two boundary layers so
 * that they run concurrently with the interior layers); ordering against tk_lattice_step_end is the
 * caller's job (events). */
int tk_lattice_step_layers_on(tk_plan *plan, int32_t zl_begin, int32_t zl_end, void *cuda_stream);
/* After every layer was updated: finish the row/column sums of the new state, swap buffers. */
int tk_lattice_step_end(tk_plan *plan);

/* ------------------------------------------------------------------------------------------
 * Fused halo push over NVLink peer memory (one process per GPU).
 * Each rank exports its two state buffers with cudaIpc (tk_lattice_ipc_export: 2 x 64-byte handles),
 * ships the handles to its z-neighbours (torch.distributed object gather), maps theirs
 * (tk_ipc_open) and attaches them (tk_lattice_peer_attach).  From then on the step kernel stores the
 * first / last owned layer of u_{t+1} into the neighbour's halo plane as it computes it, and the only
 * per-step synchronisation is a pair of stream-ordered 64-bit flags (cuStreamWriteValue64 /
 * cuStreamWaitValue64, no kernel spins): tk_lattice_signal(v) after the boundary launch,
 * tk_lattice_wait(v) before the next one.  side: 0 = lower neighbour (rank-1), 1 = upper.
 * ------------------------------------------------------------------------------------------ */
int tk_lattice_ipc_export(tk_plan *plan, void *handles_2x64);
int tk_ipc_open(int device, const void *handle_64, void **dev_ptr);
int tk_ipc_close(int device, void *dev_ptr);
/* peer_buf0/1: the neighbour's two state buffers (device pointers valid in this process);
 * peer_layers: number of layers the neighbour owns.  Pass NULL pointers to detach that side. */
int tk_lattice_peer_attach(tk_plan *plan, int32_t side, void *peer_buf0, void *peer_buf1,
                           int32_t peer_layers);
/* in_kernel = 1 (default): the step kernel pushes the boundary planes itself (fused).  0: it does not;
 * the caller moves them with tk_lattice_push_halos after the step (copy engines, off the SMs). */
int tk_lattice_set_push_mode(tk_plan *plan, int32_t in_kernel);
/* Copy the boundary planes of the CURRENT state into the attached neighbours' halos (initial fill). */
int tk_lattice_push_halos(tk_plan *plan, void *cuda_stream);
/* Stream-ordered: write `value` into both attached neighbours' arrival flags / wait until both of
 * this rank's arrival flags are >= value. */
int tk_lattice_signal(tk_plan *plan, uint64_t value, void *cuda_stream);
int tk_lattice_wait(tk_plan *plan, uint64_t value, void *cuda_stream);
/* Early signal: arm the NEXT tk_lattice_step_layers(plan, 0, all layers) launch to write `value` into the neighbours'
 * arrival flags itself, as soon as the warps of the first / last owned layer (the only ones that read this rank's halos
 * and push into the neighbours') have finished -- the boundary layer groups are scheduled first in that launch -- instead
 * of a stream-ordered write after the whole kernel.  tk_lattice_signal_pending tells whether the launch took the
 * request (0) or the caller still has to call tk_lattice_signal (1: fp32 plans, split launches, kernel variants). */
int tk_lattice_arm_signal(tk_plan *plan, uint64_t value);
int tk_lattice_signal_pending(tk_plan *plan, int32_t *pending);
/* Device addresses of this plan's two state buffers (for attaching plans of the same process). */
int tk_lattice_state_ptrs(tk_plan *plan, void **buf0, void **buf1);

/* Tiling the lattice step kernel uses (diagnostics for bench.py): cells per lane, segments per
 * CTA, rows per CTA, resident CTAs per SM.  Overridable with TK_LAT_CPL / TK_LAT_WC / TK_LAT_RCHUNK. */
int tk_lattice_tiling(const tk_plan *plan, int32_t *cells_per_lane, int32_t *segs_per_cta,
                      int32_t *rows_per_cta, int32_t *ctas_per_sm);

/* ------------------------------------------------------------------------------------------
 * Probe spectrum: |DFT| of `batch` real series of length n (physics.py:162-163, np.fft.fft of a
 * real signal; n is arbitrary, steps+1 is never a power of two) and np.argmax of each
 * (batch.py:205).  series: batch x n, spectrum: batch x n, argmax: batch (may be NULL).
 * ------------------------------------------------------------------------------------------ */
int tk_dft_abs(int device, const double *series, int64_t batch, int64_t n, double *spectrum,
               int64_t *argmax);

/* ------------------------------------------------------------------------------------------
 * Batched ensembles: n_sims independent simulations of ONE small lattice shape, each with its own
 * defect mask, kick node, coeff = (c*dt_eff)^2 and damping; one simulation per CTA, state resident
 * in shared memory.  Replaces one Python process per parameter point
 * (examples/naked_singularity_sweep.py:211-212) and the serial loop of evals/dataset.py:81-163.
 *   cell_bits  [n_masks][layers*size*size]  bits 0-3 part(q), bits 4-7 alive(q)
 *   exc_*      tagged nodes of each mask (CSR over masks): node index, inv_mass, potential
 *   corr_*     directed edge corrections of each mask (CSR over masks, <= 32 per mask)
 *   sim_*      per simulation: mask id, kicked (= probed) node, kick value, coeff, damping
 * Outputs (host): probe series [n_sims][steps+1]; |DFT| of it [n_sims][steps+1] (or NULL);
 * np.argmax of the spectrum [n_sims] (or NULL); device time of the kernel in ms (or NULL).
 * ------------------------------------------------------------------------------------------ */
typedef struct tk_ensemble_desc {
    int32_t size;
    int32_t layers;
    int64_t n_masks;
    const uint8_t *cell_bits;
    const int32_t *exc_offset; /* [n_masks+1] or NULL */
    const int32_t *exc_node;
    const double *exc_inv_mass;
    const double *exc_potential;
    const int32_t *corr_offset; /* [n_masks+1] or NULL */
    const int32_t *corr_a;
    const int32_t *corr_b;
    const double *corr_sign;
    int64_t n_sims;
    const int32_t *sim_mask;
    const int32_t *sim_kick;
    const double *sim_kick_value; /* NULL = 1.0 */
    const double *sim_coeff;
    const double *sim_damping;
    int64_t steps;
} tk_ensemble_desc;

int tk_ensemble_run(int device, const tk_ensemble_desc *desc, double *probe_series,
                    double *spectrum, int64_t *argmax, double *elapsed_ms);

/* Same, plus the spectral and time-series features of every simulation's probe series computed in the kernel's epilogue
 * (reference tetrakis_sim/evals/features.py:11-81, what tetrakis-eval generate writes per record, evals/dataset.py:146-153):
 * features[n_sims][TK_ENSEMBLE_NFEAT] =
 *   0 dominant bin, 1 spectral centroid, 2 spectral bandwidth (both in bins: divide by (steps+1)*dt_eff for the
 *   reference's frequency units), 3 peak amplitude, 4 low/high ratio, 5-9 bins of the five largest peaks (-1 = none),
 *   10-14 their magnitude relative to the maximum, 15 rms and 16 max |.| of the series, 17 sum of the half spectrum;
 * all on the non-negative-frequency half of the spectrum (features.py:18-21).  probe_series / spectrum / argmax may be
 * NULL when only the features are wanted (18 doubles per simulation instead of 2 x (steps+1)). */
#define TK_ENSEMBLE_NFEAT 18
int tk_ensemble_run_ex(int device, const tk_ensemble_desc *desc, double *probe_series, double *spectrum,
                       int64_t *argmax, double *features, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* TETRAKIS_B200_H */
