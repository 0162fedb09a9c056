/*
 * justpic_c.h -- C ABI of libjustpic_sm100a.so
 *
 * The drop-in boundary for JustPIC.jl's particle-in-cell hot path on NVIDIA
 * B200 (sm_100a).  JustPIC.jl has no FFI of its own: its seam is Julia dispatch
 * on the backend type, ending in `launch!(ka_backend(x), kernel!, ndrange, ...)`
 * (reference src/launch.jl:65-69) with allocation/conversion supplied by
 * ext/JustPICCUDAExt.jl:22-45.  Each entry point below replaces one L4 launcher
 * of the reference (file:line cited per function); julia/JustPICSM100aExt.jl
 * shows the `ccall` stubs a maintainer would add, INTEGRATION.md explains them.
 *
 * Conventions
 *  - Every array pointer is a DEVICE pointer owned by the caller (CuArray /
 *    torch tensor); the library never retains it past the call.  Only
 *    jp_grid_desc carries HOST pointers (the grid vectors are a few KB and are
 *    copied into the context once).
 *  - CellArray layout = the reference's CUDA layout (blocklength 0,
 *    ext/JustPICCUDAExt.jl:26-30): element (cell c, slot s) at c + s*C,
 *    c = i + nx*(j + ny*k) 0-based, C = nx*ny*nz.  `index` is 1 byte per slot.
 *  - Grid-node arrays are column-major (x fastest) like Julia Arrays.
 *  - All calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *    the reference's "launch then synchronize" semantics (src/launch.jl:60-69)
 *    are kept by the caller synchronising the stream.
 *  - Return 0 on success, negative jp_status on error; jp_last_error() returns
 *    a thread-local message.  Nothing throws across the boundary.
 *  - There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef JUSTPIC_C_H
#define JUSTPIC_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JP_MAX_ARGS 16   /* particle fields carried by move/inject/clean/halo in one call */
#define JP_MAX_SLOTS 64  /* max_xcell served by the occupancy-word kernels (the tuned path) */
#define JP_MAX_SLOTS_WIDE 1024  /* largest max_xcell accepted: above JP_MAX_SLOTS the per-slot kernels run in 64-slot
                                  chunks and move / inject take literal per-cell kernels (same results, not tuned) */
#define JP_MAX_PHASES 32

typedef enum {
    JP_OK = 0,
    JP_ERR_INVALID = -1,     /* bad argument (dims, null pointer, alpha out of (0,1), ...) */
    JP_ERR_CUDA = -2,        /* CUDA runtime error (message holds cudaGetErrorString) */
    JP_ERR_UNSUPPORTED = -3  /* e.g. S > JP_MAX_SLOTS_WIDE, nargs > JP_MAX_ARGS */
} jp_status;

typedef enum { JP_EULER = 0, JP_RK2 = 1, JP_RK4 = 2 } jp_scheme;
typedef enum { JP_INTERP_LINEAR = 0, JP_INTERP_LINP = 1, JP_INTERP_MQS = 2 } jp_interp;
/* JP_OPT_ADVECT_AFFINE (0/1, default 1): let the tiled advection kernel regenerate grid coordinates as
 * fma(i, dx, x0) when jp_ctx_create verified that this reproduces EVERY stored entry bit for bit
 * (results are identical either way; 0 forces the table look-ups, used by the parity tests). */
/* JP_OPT_MOVE_POLICY (default JP_MOVE_POLICY_REFERENCE): which free slot a migrant takes.
 *   REFERENCE: the reference's rule, bit for bit -- first free slot >= a cursor that is carried over from one
 *     migrant of a source cell to the next, across destination cells (src/Particles/move_safe.jl:114-118).
 *     It spreads the particles of a cell over all max_xcell slots (slot planes ~50 % full).
 *   COMPACT: NOT reference behaviour, opt-in -- the search starts at slot 0 for every migrant (the one
 *     line `starting_point = free_idx` dropped).  Same particles in the same cells (up to fewer drops in
 *     over-full cells), but in the lowest free slots: slot planes stay dense, so every streaming kernel
 *     moves fewer dead-slot sectors.  Slot positions -- hence masks and summation order -- differ from the
 *     reference's; checked against the oracle run with the same rule.
 *   DENSE: NOT reference behaviour, opt-in -- "vacate everything, then place": all cells give up their leavers' slots
 *     first; then the migrants are placed in the reference's order (3^N colours, source cells, slot order), each in the
 *     LOWEST free slot of its destination; dropped when the destination is full.  A migrant thus finds the slots its
 *     destination vacates in the same call (under the other two rules only when the destination's colour came first),
 *     so the cells stay packed towards slot 0.  Planned path only (JP_MOVE_AUTO, max_xcell <= 64: jp_move refuses
 *     otherwise); a call the planner declines (a particle exactly on a cell face, a move of more than one cell --
 *     jp_last_move_path says so) runs the in-place sweeps with the REFERENCE rule.  Oracle twin: jpo_set_move_policy(2). */
/* JP_OPT_ADVECT_CLASSIFY (0/1, default 0): advection -> move hand-off.  With 1, jp_advect (tiled kernel:
 *   standard staggering) also classifies every new position for the following move_particles! -- the
 *   same comparisons jp_move makes, on the value being stored -- and writes the move plan's per-cell
 *   occupancy / leave / destination-code words into the library workspace; the next jp_move on the SAME
 *   coordinate / index arrays starts from those words instead of re-reading the coordinates (results
 *   bit-identical to option 0).  The words are dropped by every library call that changes particles
 *   (init, inject, clean, another advect, move); planes rewritten by jp_halo_unpack are re-classified
 *   from the coordinates.  The caller must not modify coordinates or the index mask between the two
 *   calls by other means -- the reference's time loops never do (advection! -> update_halo! ->
 *   move_particles!), but the library cannot see such writes, hence opt-in.
 * JP_OPT_LAST_CLASSIFY (jp_get_option only): 1 if the last planned jp_move used the hand-off bytes. */
/* JP_OPT_MOVE_INTERP (0/1, default 0): move -> interpolation hand-off.  With 1, the last streaming pass of jp_move (which
 *   already touches nearly every sector of every particle array) also accumulates, for each cell's FINAL content and in slot
 *   order, (a) the per-cell partial sums of the two-pass particle2grid! of the particle field registered with
 *   jp_move_interp_fields and (b) the centre phase ratios of the registered phase field -- the same arithmetic in the same
 *   order as jp_particle2grid's cell pass / jp_phase_ratios_center, results bit-identical to option 0.  The next
 *   jp_particle2grid(F, Fp) with that Fp then only runs its node pass and the next jp_phase_ratios_center(ratios, phases, K)
 *   copies the ratios out; neither reads the particles again.  Dropped by every library call that changes particles or writes
 *   a particle field (init, advect, inject, clean, force_injection, halo unpack, grid2particle, ...); the caller must not
 *   modify the particle arrays or the two fields between jp_move and the consumers by other means -- the reference's time
 *   loops never do (move_particles! -> particle2grid! -> phase ratios) -- hence opt-in.  Needs max_xcell <= 64, a two-pass
 *   particle2grid mode, <= 4 phases, and containers whose dead slots hold NaN coordinates (everything init_particles /
 *   move_particles! / inject_particles! produce): liveness is the occupancy word, not phase_ratios_center!'s isnan(px) probe.
 * JP_OPT_GRAPH_STEP_OFFSET (value >= 0; default 0): CUDA graphs.  Every entry point may be called on a stream that is being captured
 *   (cudaStreamBeginCapture, CUDA.jl's @captured, torch.cuda.graph): nothing on the hot path reads the device back, jp_move skips its
 *   asynchronous staging-size feedback, and what cannot be captured -- (re)allocating a workspace, the very first jp_move of a context --
 *   returns JP_ERR_INVALID saying so (run the same calls once eagerly first).  Replays run with the arguments of the captured call; the
 *   one argument that must change from call to call, the `step` of jp_inject / jp_inject_phase, is offset on the device by this word:
 *   a captured inject call increments it when it has run, so replay i injects with step + i, like the eager loop.  Set it back to 0
 *   (synchronous) when going back to eager calls or capturing a new graph; jp_get_option reads it.
 * JP_OPT_LAST_INTERP (jp_get_option only): bit 0 / bit 1 = the last jp_particle2grid / jp_phase_ratios_center used the hand-off. */
typedef enum { JP_OPT_P2G_MODE = 1, JP_OPT_MOVE_MODE = 2, JP_OPT_ADVECT_AFFINE = 3, JP_OPT_MOVE_POLICY = 4,
               JP_OPT_ADVECT_CLASSIFY = 5, JP_OPT_LAST_CLASSIFY = 6, JP_OPT_MOVE_INTERP = 7, JP_OPT_LAST_INTERP = 8, JP_OPT_PROFILE = 9, JP_OPT_GRAPH_STEP_OFFSET = 10 } jp_option;
typedef enum { JP_MOVE_POLICY_REFERENCE = 0, JP_MOVE_POLICY_COMPACT = 1, JP_MOVE_POLICY_DENSE = 2 } jp_move_policy;
typedef enum { JP_MOVE_AUTO = 0, JP_MOVE_DIRECT = 1 } jp_move_mode;
typedef enum { JP_P2G_EXACT = 0, JP_P2G_TWOPASS = 1, JP_P2G_TWOPASS_FASTW = 2 } jp_p2g_mode;

/* Grid description (HOST pointers).  Mirrors the grid part of the reference's
 * `Particles` struct (src/particles.jl:17-46): xvi, xci, xi_vel, and whether
 * `di` holds scalar spacings (range grids, particles_utils.jl:137-140) or
 * diff() vectors (array grids, particles_utils.jl:76-79). */
typedef struct {
    int32_t ndim;              /* 2 or 3 */
    int32_t n[3];              /* cells per dimension (n[2] ignored in 2D) */
    int32_t S;                 /* slots per cell = max_xcell */
    int32_t uniform;           /* 1: spacing = x[1]-x[0] everywhere; 0: x[i+1]-x[i] */
    const double *xv[3];       /* vertex coordinates, n[d]+1 */
    const double *xc[3];       /* centre coordinates, n[d] */
    const double *xvel[3][3];  /* xvel[comp][dim]: staggered velocity grid vectors */
    int32_t nvel[3][3];        /* their lengths (n, n+1 or n+2) */
} jp_grid_desc;

/* Particle storage (DEVICE pointers): coords[d] and index are [C*S]. */
typedef struct {
    double *coords[3];
    uint8_t *index;
} jp_particles;

/* Opaque context: device copy of the grid + derived tables + workspace. */
typedef struct jp_ctx jp_ctx;

/* Create/destroy a context on CUDA device `device`.  One per Particles object
 * (the Julia shim creates it in init_particles and caches it). */
int  jp_ctx_create(const jp_grid_desc *grid, int device, jp_ctx **out);
void jp_ctx_destroy(jp_ctx *ctx);
const char *jp_last_error(void);
int  jp_version(void);
int  jp_set_option(jp_ctx *ctx, int32_t option, int32_t value);
int  jp_get_option(const jp_ctx *ctx, int32_t option, int32_t *value);   /* effective value (e.g. AFFINE = option && detected) */

/* init_particles(backend, nxcell, max_xcell, min_xcell, xi_vel...)
 * (src/Particles/particles_utils.jl:108-166, kernel fill_coords_index! :168-194).
 * Fills coords with NaN / index with 0, then seeds ceil(nxcell/2^N) particles per
 * quadrant with Philox4x32-10 keyed (seed, cell, slot). */
int jp_init_particles(jp_ctx *ctx, const jp_particles *p, int32_t nxcell, uint64_t seed, void *stream);

/* advection!(particles, method, V, dt)
 * (src/Particles/Advection/advection.jl:21-91; Euler.jl, RK2.jl, RK4.jl).
 * V[comp] are the staggered velocity arrays with extents nvel[comp][0..ndim). */
int jp_advect(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha,
              const double *const *V, double dt, void *stream);

/* advection! in two launches, for overlapping update_cell_halo! (src/CellArrays/ImplicitGlobalGrid.jl:36-41, called between
 * advection! and move_particles! in scripts/temperature_advection3D_MPI.jl:83-91) with the bulk of the advection:
 * JP_REGION_SHELL advects every brick of cells that holds one of the two outermost cell layers -- everything the exchange
 * reads (layers 1, n-2) or rewrites (layers 0, n-1) --, JP_REGION_INTERIOR the remaining bricks; the pair equals one
 * jp_advect (bit-identical results, same hand-off to jp_move).  The caller runs the exchange on another stream after the
 * SHELL call and joins before jp_move.  JP_REGION_ALL = jp_advect. */
typedef enum { JP_REGION_ALL = 0, JP_REGION_SHELL = 1, JP_REGION_INTERIOR = 2 } jp_region;
int jp_advect_region(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt,
                     int32_t region, void *stream);

/* advection_LinP!(particles, method, V, dt) (src/Particles/Advection/advection_LinP.jl:12-391) and
 * advection_MQS!(particles, method, V, dt) (advection_MQS.jl:16-124, src/Interpolations/MQS.jl):
 * same integrators, velocity reconstructed with the LinP / MQS interpolant where the interpolation
 * cell is interior (linear elsewhere).  interp = JP_INTERP_LINP / JP_INTERP_MQS (LINEAR = jp_advect). */
int jp_advect_interp(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha,
                     const double *const *V, double dt, int32_t interp, void *stream);

/* move_particles!(particles, args) (src/Particles/move_safe.jl:21-125).
 * Slot assignment bit-exact with the reference's 3^N colour sweeps (for displacements
 * <= 1 cell; larger ones are racy in the reference itself).  JP_MOVE_AUTO plans the
 * sweeps on per-cell occupancy words and moves payloads in two streaming passes, and
 * falls back to JP_MOVE_DIRECT (literal sweeps on the particle arrays) when a particle
 * sits exactly on a cell face.  Asynchronous on `stream`: the fallback is decided on the device (the kernels of
 * the path not taken return at once); only the very first planned call on a context blocks, once, to size its staging buffer. */
int jp_move(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs, void *stream);
/* JP_OPT_MOVE_INTERP: which particle field the following particle2grid!(F, Fp, particles) will interpolate and which field /
 * how many phases the following phase_ratios_center!(phase_ratios, particles, phases) will use (either may be NULL).  Both must
 * be among the `args` of jp_move to take effect.  Sticky until changed. */
int jp_move_interp_fields(jp_ctx *ctx, const double *Fp, const double *phases, int32_t K);
/* Both hand-offs are keyed on the coordinate / index (and field) POINTERS: a write to those arrays that does not go through a jp_*
 * entry point (a host-side boundary fix-up, a torch / CUDA.jl kernel of the caller) between the producing and the consuming call is
 * invisible to the library.  Call this after such a write: the consumers then work from the arrays again. */
int jp_invalidate_handoffs(jp_ctx *ctx);
/* JP_OPT_PROFILE = 1: jp_move (planned path) records CUDA events between its stages on the caller's stream (no
 * synchronisation); jp_profile_read returns the mean duration in ms of {classify, plan (3^N launches), finalize + scan,
 * gather, scatter (+ interpolation hand-off)} over the calls since the last read (at most the last 32) and synchronises the
 * device.  This is how bench.py times the dominant kernel inside the move phase. */
int jp_profile_read(jp_ctx *ctx, double out_ms[5], int32_t *ncalls);
/* Counters of the last jp_move on this context: {moved, dropped (destination
 * full), deleted (left the domain)}.  Synchronises `stream`. */
int jp_move_stats(jp_ctx *ctx, int64_t out[3], void *stream);
/* bits 0-7: 0 = the last jp_move took the plan/gather/scatter path, 1 = direct sweeps; bits 8+: why (1 move of more than one
 * cell, 2 / 4 particle on a face of its own / destination cell, 8 staging buffer too small).  Synchronises that jp_move's stream. */
int jp_last_move_path(const jp_ctx *ctx);

/* inject_particles!(particles, args) (src/Particles/injection.jl:19-131).
 * RNG: Philox4x32-10 keyed (seed, step, cell, slot). */
int jp_inject(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs,
              int32_t min_xcell, uint64_t seed, uint32_t step, void *stream);
/* Number of particles injected by the last jp_inject.  Synchronises `stream`. */
/* inject_particles_phase!(particles, particles_phases, args, fields) (src/Particles/injection.jl:146-325):
 * every quadrant of every cell is examined (2^N colour sweeps as jp_inject); a new particle takes the
 * PHASE of its nearest live neighbour (index_min_distance, :330-393) and args[j] is interpolated from the
 * grid field fields[j] at the new position, clamped to the extrema of the interpolation stencil:
 * field_kind[j] = 1: centre field (n per dim; shifted/clamped centre cell, :295-305), 0: vertex field
 * (n+1 per dim; the storage cell, :307-311).  Counter-based RNG keyed (seed, purpose 2, step, cell, slot). */
int jp_inject_phase(jp_ctx *ctx, const jp_particles *p, double *phases, double *const *args, const double *const *fields,
                    const int32_t *field_kind, int32_t nargs, int32_t min_xcell, uint64_t seed, uint32_t step, void *stream);
int jp_inject_stats(jp_ctx *ctx, int64_t *out, void *stream);

/* force_injection!(particles, p_new, fields, values) (src/Particles/forced_injection.jl:16-79; reference tests
 * test/test_2D.jl:301-384, test/test_3D.jl:279-333).  pnew[d]: [C*S] coordinate component d of the caller's candidate
 * points, element (cell c, entry k) at c + k*C (the reference's p_new[I..., k]).  A cell injects iff its FIRST entry is
 * not NaN in component 0 ("NaN marks empty input", :9, :36); then every free slot ip of the cell takes entry ip (the
 * reference's counter c advances with the slot loop, :38-42), its mask is set and fields[j][slot] = values[j]. */
int jp_force_injection(jp_ctx *ctx, const jp_particles *p, const double *const *pnew, double *const *fields, const double *values,
                       int32_t nfields, void *stream);

/* clean_particles!(particles, grid, args) (src/Particles/move_safe.jl:289-320). */
int jp_clean(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs, void *stream);

/* grid2particle!(Fp, F, particles) (src/Interpolations/grid_to_particle.jl:26-82).
 * F: vertex field (n+1 per dim).  Fp: [C*S]. */
int jp_grid2particle(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *F, void *stream);

/* centroid2particle!(Fp, F, particles) (src/Interpolations/centroid_to_particle.jl:13-46).
 * Fc: centre field (n per dim). */
int jp_centroid2particle(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *Fc, void *stream);

/* particle2grid!(F, Fp, particles) (src/Interpolations/particle_to_grid.jl:23-151).
 * Two modes (jp_set_option(ctx, JP_OPT_P2G_MODE, ...)):
 *  JP_P2G_TWOPASS: cell-centric partial sums + node-centric gather, no
 *    atomics, fixed summation order; every particle is read once.  Same weights and
 *    terms as the reference, different association: agrees to a few ulp (stated
 *    tolerance 1e-12 relative), deterministic run to run.
 *  JP_P2G_TWOPASS_FASTW (default): as TWOPASS with the weight evaluated as 1/sum(d^2) instead of the
 *    reference's inv(sqrt(sum(d^2))^2) (differs by <= 2 ulp per weight; same 1e-12 tolerance).
 *  JP_P2G_EXACT: one thread per node, the reference's single running sum in its
 *    (k, j, i, slot) order: bit-exact with the reference, reads every particle 2^N times. */
int jp_particle2grid(jp_ctx *ctx, const jp_particles *p, double *F, const double *Fp, void *stream);

/* grid2particle_flip!(Fp, xvi, F, F0, particles; alpha) (src/Interpolations/grid_to_particle.jl:125-173):
 * Fp <- muladd(F_pic, alpha, (Fp + (F_pic - F0_pic)) * (1 - alpha)), F_pic / F0_pic = vertex fields F / F0
 * interpolated with the spacing grid_size(xvi) (the minimum spacing per dimension, utils.jl:61-63). */
int jp_grid2particle_flip(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *F, const double *F0,
                          double alpha, void *stream);

/* subgrid_diffusion!(pT, T_grid, dT_grid, subgrid_arrays, particles, dt; d) and subgrid_diffusion_centroid!
 * (src/Physics/subgrid_diffusion.jl:55-143).  pT0 / pdT / dt0 / dT_subgrid are the fields of
 * SubgridDiffusionCellArrays (:17-33): particle CellArrays [C*S] and a vertex (centroid = 0) or centre
 * (centroid = 1) grid array.  dT_grid is read at I + 1 (one ghost node on the low side, :137), its extents
 * are dT_extents[0..ndim).  Same sequence of operations as the reference's seven launches, fused into two
 * particle passes around particle2grid!/particle2centroid!; uses exp(), so the stated 1e-12 applies. */
int jp_subgrid_diffusion(jp_ctx *ctx, const jp_particles *p, double *pT, const double *T_grid, const double *dT_grid,
                         const int32_t *dT_extents, double *pT0, double *pdT, const double *dt0, double *dT_subgrid,
                         double dt, double d, int32_t centroid, void *stream);

/* particle2centroid!(F, Fp, particles) (src/Interpolations/particle_to_grid_centroid.jl:10-99). */
int jp_particle2centroid(jp_ctx *ctx, const jp_particles *p, double *Fc, const double *Fp, void *stream);

/* phase_ratios_center!(phase_ratios, particles, phases) (src/PhaseRatios/centers.jl:3-30).
 * ratios: CellArray [C*K], element (cell c, phase k) at c + k*C.  phases: [C*S] fp64 ids 1..K. */
int jp_phase_ratios_center(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases,
                           int32_t K, void *stream);

/* phase_ratios_vertex!(phase_ratios, particles, phases) (src/PhaseRatios/vertices.jl:4-107).
 * ratios: CellArray over the vertex grid (n+1 per dim), element (node, k) at node + k*NN.  Every
 * particle of the 2^N adjacent cells within (strictly less than) half a cell of the vertex
 * contributes; a vertex with none keeps the reference's NaN (0 * inv(0)). */
int jp_phase_ratios_vertex(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases,
                           int32_t K, void *stream);

/* phase_ratios_face!(phase_face, particles, phases, dimension) (src/PhaseRatios/midpoints.jl:3-82).
 * dim = 0/1/2 for :x/:y/:z.  ratios: CellArray over the face grid n + e_dim (the Vx/Vy/Vz fields of
 * PhaseRatios, constructors.jl:41-43), element (node, k) at node + k*NF; NaN (no particle in range) -> 0. */
int jp_phase_ratios_face(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases,
                         int32_t K, int32_t dim, void *stream);

/* phase_ratios_midpoint!(phase_midpoint, particles, phases, dimension) (src/PhaseRatios/midpoints.jl:115-242),
 * 3-D only.  plane = 0/1/2 for :xy/:yz/:xz, midpoint grid n + offsets with offsets (1,1,0)/(0,1,1)/(1,0,1)
 * (the xy/yz/xz fields of PhaseRatios, constructors.jl:44-46); NaN -> 0.  The boundary branch is the
 * reference's, quirks included (listed in DESIGN.md). */
int jp_phase_ratios_midpoint(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases,
                             int32_t K, int32_t plane, void *stream);

/* update_phase_ratios!(phase_ratios, particles, phases) (src/PhaseRatios/utils.jl:15-41): centre, vertex, the
 * ndim face fields (Vx, Vy[, Vz]) and, in 3-D, the midpoint fields in the order xy, yz, xz.
 * JP_PHASE_LITERAL: the reference's sequence of kernels, bit-exact.  JP_PHASE_FUSED: ONE pass over the particles
 * accumulating per-cell partial sums for every node a cell touches + node-centric gathers in the reference's
 * cell order (the low-boundary nodes still come from the literal boundary branches): same terms, associated
 * as (cell sums) + ..., within the stated 1e-12; centre ratios stay bit-exact.  FUSED needs K <= 4 (else literal). */
typedef enum { JP_PHASE_LITERAL = 0, JP_PHASE_FUSED = 1 } jp_phase_mode;
int jp_update_phase_ratios(jp_ctx *ctx, const jp_particles *p, const double *phases, int32_t K, double *center, double *vertex,
                           double *const *faces, double *const *midpoints, int32_t mode, void *stream);

/* update_cell_halo! building blocks (src/CellArrays/ImplicitGlobalGrid.jl:36-41):
 * gather / scatter one cell-plane (all S slots) of every listed CellArray into /
 * from one contiguous device buffer; the transport between ranks (NCCL
 * send/recv) is done by the host layer.  Buffer layout: for each fp64 array in
 * order, then the index bytes; within an array slot-major, then the plane's
 * cells in memory order.  jp_halo_plane_bytes returns the buffer size. */
int64_t jp_halo_plane_bytes(const jp_ctx *ctx, int32_t dim, int32_t narrays);
int jp_halo_pack(jp_ctx *ctx, int32_t dim, int32_t plane, double *const *arrays, int32_t narrays,
                 const uint8_t *index, void *buf, void *stream);
int jp_halo_unpack(jp_ctx *ctx, int32_t dim, int32_t plane, double *const *arrays, int32_t narrays,
                   uint8_t *index, const void *buf, void *stream);

/* update_cell_halo!(particles.coords..., args..., particles.index) (src/CellArrays/ImplicitGlobalGrid.jl:36-41 =
 * ImplicitGlobalGrid.update_halo! on every CellArray; scripts/temperature_advection3D_MPI.jl:86) as ONE library call: for each
 * dimension x -> y -> z the planes 2 / n-1 (1-based; overlap 2, halo width 1) of all listed CellArrays and the index mask are
 * packed into one message per face, exchanged with ncclSend / ncclRecv inside one ncclGroup and unpacked into the neighbours'
 * planes n / 1, so edges and corners propagate through the sequential dimensions.  comm: an ncclComm_t (from jp_comm_init, or
 * the host's own NCCL binding -- NCCL.jl's communicator handle); nbr[2*d], nbr[2*d+1]: rank of the left / right neighbour
 * along dimension d in that communicator, -1 at a non-periodic boundary; a rank that is its own neighbour (periodic, one
 * rank along d) wraps around locally and needs no communicator.  Asynchronous on `stream`. */
int jp_halo_exchange(jp_ctx *ctx, void *comm, const int32_t *nbr, double *const *arrays, int32_t narrays, uint8_t *index,
                     void *stream);
/* update_halo!(A) for a plain grid array A (a staggered velocity component, a vertex field ...) with extents ext[0..ndim) on
 * the same decomposition: overlap = 2 + ext[d] - n[d] as ImplicitGlobalGrid computes it for staggered arrays.  This is the
 * velocity-ghost-layer exchange when V comes from a solver rather than a formula. */
int jp_halo_exchange_grid(jp_ctx *ctx, void *comm, const int32_t *nbr, double *A, const int32_t *ext, void *stream);
/* NCCL plumbing for hosts without their own binding.  libnccl.so.2 is dlopen'ed on first use (override: JP_NCCL_LIB);
 * jp_comm_unique_id on one rank -> broadcast the 128 bytes by any means -> jp_comm_init on every rank (collective). */
int jp_comm_unique_id(void *id128);
int jp_comm_init(const void *id128, int32_t nranks, int32_t rank, int32_t device, void **comm_out);
int jp_comm_destroy(void *comm);
/* dt = min(dx / MPI.Allreduce(maximum(abs(V)), MPI.MAX)) of the reference's MPI scripts (temperature_advection3D_MPI.jl:71):
 * in-place max over the ranks of n device doubles. */
int jp_allreduce_max(void *comm, double *buf, int32_t n, void *stream);

/* Array(CA) / Array(T, CA) and CuArray(CA) / CuArray(T, CA) for a CellArray
 * (src/CellArrays/conversion.jl:19-43, ext/JustPICCUDAExt.jl:166-179; checkpoints, test/test_save_load.jl:120-173):
 * the reference's permutedims(CA.data, (3, 2, 1)) between the device layout data[C, S, 1] (element
 * (cell c, component s) at c + s*C) and the host layout data[1, S, C] (at s + c*S), with the element
 * conversion of the typed forms (Float64 <-> Float32 by IEEE round-to-nearest as Julia's convert; Bool
 * stays Bool).  src and dst are both DEVICE pointers of ncells*ncomp elements and must not alias: the
 * shim copies dst to the host (JP_LAYOUT_TO_HOST) or uploads src first (JP_LAYOUT_TO_DEVICE).
 * ncells = number of cells of that CellArray (n for particle fields / centres, n+1 per dim for vertex
 * ratios, ...), ncomp = entries per cell (max_xcell, nphases, ...).  ctx may be NULL (a bare CellArray has
 * no Particles): the launch then goes to the caller's current device. */
typedef enum { JP_F64 = 0, JP_F32 = 1, JP_BOOL = 2 } jp_dtype;
typedef enum { JP_LAYOUT_TO_HOST = 0, JP_LAYOUT_TO_DEVICE = 1 } jp_layout_direction;
int jp_cellarray_permute(jp_ctx *ctx, const void *src, int32_t src_type, void *dst, int32_t dst_type,
                         int64_t ncells, int32_t ncomp, int32_t direction, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* JUSTPIC_C_H */
