/* veritas_b200 — C ABI of the B200-native Vlasov advance (drop-in boundary for Libbum/Veritas).
 *
 * The reference has no FFI: its "API" is the C++ class surface used by veritas.cpp.
 * Each entry point below names the reference member function(s) whose body it replaces, so the
 * reference classes become thin shells around these calls (see INTEGRATION.md and the host classes
 * in veritas_b200/host/, which do exactly that).
 *
 * Conventions: every call returns 0 on success or a negative code; vrt_last_error() gives the text.
 * Nothing throws, nothing calls exit().  One host thread per context; all device work of a context
 * is issued on one CUDA stream.  Device memory is owned by the context, host buffers by the caller.
 * All arithmetic is fp64.  There is no CPU fallback: without a CUDA device vrt_create() fails.
 *
 * Patch arrays exchanged with the host use the reference's padded layout (2 ghost cells per side,
 * p fastest): index (n_p+4)*(i+2)+2+j for cell (i,j)  (Rectangle.hpp:105-108), one plane per state.
 */
#ifndef VERITAS_B200_H
#define VERITAS_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vrt_ctx vrt_ctx;

enum { VRT_OK = 0, VRT_ERR_ARG = -1, VRT_ERR_CUDA = -2, VRT_ERR_STATE = -3, VRT_ERR_NCCL = -4, VRT_ERR_NOMEM = -5 };

/* field selectors for vrt_field_upload/download: the six transverse arrays of EMFieldSolver
 * (EMSolver.hpp:18), 8 time slots each (EMSolver.hpp:51-53) */
enum { VRT_BY = 0, VRT_BZ = 1, VRT_EY = 2, VRT_EZ = 3, VRT_AY = 4, VRT_AZ = 5 };
/* 1-D array selectors for vrt_get_1d/vrt_set_1d */
enum { VRT_PHI = 0, VRT_CHARGE = 1, VRT_J = 2, VRT_A_SQUARED = 3, VRT_NEUTRALIZATION = 4, VRT_EFIELD = 5,
       VRT_CHARGES0 = 16 /* + species index: per-species charge (EMSolver.hpp:19 `charges`) */ };
/* scalar selectors */
enum { VRT_EX0 = 0, VRT_TIME = 1 };
/* Vlasov kernel selection */
enum { VRT_PATH_AUTO = 0, VRT_PATH_SPLIT = 1 /* sub-step kernels with ghost syncs, any hierarchy */,
       VRT_PATH_FUSED = 2 /* one streaming pass per stage; single-level full-domain patches / x-slabs */ };

/* work-plane selectors for vrt_patch_download_plane (Rectangle.hpp:7: FxH, FpH, FxL, FpL, FxDS, FpDS, Rp, Rm, Cx, Cp, ex, ep, fx, fp) */
enum { VRT_PLANE_FXH = 0, VRT_PLANE_FPH, VRT_PLANE_FXL, VRT_PLANE_FPL, VRT_PLANE_FXDS, VRT_PLANE_FPDS, VRT_PLANE_RP, VRT_PLANE_RM,
       VRT_PLANE_CX, VRT_PLANE_CP, VRT_PLANE_EX, VRT_PLANE_EP, VRT_PLANE_FX, VRT_PLANE_FP };

/* Rectangle ctor arguments (Rectangle.hpp:23): depth 0 = finest level. */
typedef struct {
    int depth, x_pos, p_pos, n_x, n_p;
    int up, down, left, right;
} vrt_patch_desc;

/* ---- lifecycle ------------------------------------------------------------------------------ */
/* Library-wide error text for failures that have no context yet (vrt_create). */
const char* vrt_global_error(void);
const char* vrt_last_error(const vrt_ctx* ctx);
const char* vrt_version(void);
/* SolverManager::SolverManager (SolverManager.cpp:8-26): one context = meshes + field solver on one GPU. */
int vrt_create(vrt_ctx** out, int device, int n_species);
int vrt_destroy(vrt_ctx* ctx);
int vrt_sync(vrt_ctx* ctx);
/* cudaStream_t of the context, as an opaque pointer (for callers that time with CUDA events). */
void* vrt_stream(vrt_ctx* ctx);

/* ---- configuration (Settings.cpp:5-58, EMSolver.cpp:6-27) ----------------------------------- */
int vrt_set_grid(vrt_ctx* ctx, int x_size_finest, double dx_finest, int n_prepad, int n_postpad,
                 int refinement_ratio, int max_depth);
int vrt_set_species(vrt_ctx* ctx, int s, double mass, double charge, double pmin, double dp_finest);
/* Mesh::promoteHierarchyToMesh (Mesh.cpp:794-875): (re)create the patches of species s.  Connectivity
 * (Rectangle::CalculateConnectivitySame / FromFiner, Rectangle.cpp:671-864) is derived from the descriptors; it depends
 * on the order of the patches inside a level, which must be the order of Level::rectangles.  Patches of refined levels
 * (depth < max_depth) must start on multiples of the refinement ratio.  Patch data are zero after this call. */
int vrt_set_hierarchy(vrt_ctx* ctx, int s, int n_patches, const vrt_patch_desc* patches);
int vrt_set_path(vrt_ctx* ctx, int path);
int vrt_get_path(vrt_ctx* ctx, int s);
/* Regrid without the host round trip.  Mesh::promoteHierarchyToMesh(false) (Mesh.cpp:794-875):
 * replaces the hierarchy of species s (split path) by `patches` and fills the new patches from the old ones on the device —
 * Mesh::InterMeshDataTransfer (Mesh.cpp:116-130): GetDataFromCoarseLevelRectangle (Rectangle.cpp:892-918),
 * GetDataFromSameLevelRectangle (920-941), GetDataFromCoarseNewLevelRectangle (1100-1128), states 0 and 1.  Cells no
 * old or new coarser patch reaches stay 0, as in the reference.  Follow with vrt_push_data and vrt_commit_state. */
int vrt_regrid(vrt_ctx* ctx, int s, int n_patches, const vrt_patch_desc* patches);
/* Rectangle::ErrorEstimate(i, j) > refinementCriteria (Rectangle.hpp:128-130, errorWeights of Rectangle.cpp:68-74) for the
 * n_x x n_p interior cells of a patch (p fast), one byte per cell.  Rectangle::getError (Rectangle.cpp:866-890) applies its
 * margins and ORs Settings::RefinementOverride on the host. */
int vrt_error_flags(vrt_ctx* ctx, int s, int patch, const double weights[5], double criteria, unsigned char* flags_host);

/* ---- data movement (init, regrid, output) --------------------------------------------------- */
/* state: 0 = f^n, 1 = current stage value, 2 = low-order predictor (Rectangle.hpp:96-98). */
int vrt_patch_upload_f(vrt_ctx* ctx, int s, int patch, int state, const double* host_padded);
int vrt_patch_download_f(vrt_ctx* ctx, int s, int patch, int state, double* host_padded);
/* One work plane of a patch in the padded layout (split path; diagnostics and parity tests). slot: RK slot for FxH/FpH. */
int vrt_patch_download_plane(vrt_ctx* ctx, int s, int patch, int which, int slot, double* host_padded);
/* Rectangle::FCTTimeStep(…,3) applied to every patch (Mesh.cpp:869-874): f0 := f1 on the padded array. */
int vrt_commit_state(vrt_ctx* ctx, int s);
int vrt_field_upload(vrt_ctx* ctx, int which, int slot, const double* host);   /* M = x_size+pads doubles */
int vrt_field_download(vrt_ctx* ctx, int which, int slot, double* host);
int vrt_set_1d(vrt_ctx* ctx, int which, const double* host);
int vrt_get_1d(vrt_ctx* ctx, int which, double* host);   /* VRT_A_SQUARED: x_size+1, others x_size */
int vrt_set_scalar(vrt_ctx* ctx, int which, double v);
int vrt_get_scalar(vrt_ctx* ctx, int which, double* v);

/* ---- the hot path --------------------------------------------------------------------------- */
/* EMFieldSolver::AssembleRhoAndJ (EMSolver.cpp:104-122) -> Mesh::InterpolateRhoAndJToFinestMesh (Mesh.cpp:52-56)
 * -> Level::CollectRhoAndJ (Level.cpp:42-62) -> Rectangle::CalculateRhoAndJ (Rectangle.cpp:157-282). */
int vrt_moments(vrt_ctx* ctx);
/* Mesh::InterpolateRhoAndJToFinestMesh(charge, J) (Mesh.cpp:52-56) for species s: ADDS the species' charge and current on the finest
 * grid to the two host arrays (x_size_finest doubles each). */
int vrt_moments_species(vrt_ctx* ctx, int s, double* charge_host, double* j_host);
/* Rectangle::CalculateRhoAndJ (Rectangle.cpp:157-282) of one patch: its chargeR and currentR, n_x * r^depth doubles each (the
 * patch's columns on the finest x grid), as Level::CollectRhoAndJ (Level.cpp:42-62) sums them into the level's arrays.  Recomputes
 * the species' moments on the device and leaves the assembled charge / J (all species) as vrt_moments does. */
int vrt_patch_moments(vrt_ctx* ctx, int s, int patch, double* charge_r_host, double* current_r_host);
/* EMFieldSolver::EnforceChargeNeutralization (EMSolver.cpp:621-629). */
int vrt_enforce_neutralization(vrt_ctx* ctx);
/* EMFieldSolver::UpdatePotential (EMSolver.cpp:156-192): periodic 4th-order Poisson + Ex0 update.  The
 * reference's dense LU (EMSolver.cpp:28-86) is replaced by an O(N) direct solve of the same linear system. */
int vrt_poisson(vrt_ctx* ctx);
/* Mesh::Advance(dt, step) (Mesh.cpp:64-89) for species s. */
int vrt_vlasov_stage(vrt_ctx* ctx, int s, double dt, int step);
/* Level::FCTTimeStep(dt, step, subStep) (Level.cpp:12-17 -> Rectangle.cpp:1255-1623); level = depth. */
int vrt_vlasov_substep(vrt_ctx* ctx, int s, int depth, double dt, int step, int substep);
/* Mesh::PushData(val) (Mesh.cpp:91-106) and Mesh::PushBoundaryC() (Mesh.cpp:904-917). */
int vrt_push_data(vrt_ctx* ctx, int s, int val);
int vrt_push_boundary_c(vrt_ctx* ctx, int s);
/* Level::PushData(updateType, val) (Level.cpp:88-126) on the level of the given depth: 0 UpdateInterriorPoints, 1 UpdateSameLevelBoundaries,
 * 2 UpdateDifferentLevelBoundaries, 3 UpdateCornerPoints, 4 CalculateSameBoundaryC, 5 CalculateDifferentBoundaryC, 6 UpdateSameBoundaryC.
 * depth = -1 serves every level in one launch; valid for the passes that pair patches of one level only (1, 4, 6). */
int vrt_level_push(vrt_ctx* ctx, int s, int depth, int update_type, int val);
/* EMFieldSolver::RGKStep(step, dt) (EMSolver.cpp:194-202) = RGKCalculateRHS + RGKUpdateIntermediateSolution +
 * InterpolateToFaces; by0/bz0 are the host-evaluated Settings::GetBY/GetBZ(0, time) (EMSolver.cpp:501-502). */
int vrt_field_stage(vrt_ctx* ctx, int step, double dt, double by0, double bz0);
/* EMFieldSolver::EstimateCFLBound (EMSolver.cpp:631-664). */
int vrt_cfl_bound(vrt_ctx* ctx, double* out);
/* Settings::UpdateTime(step, dt) (Settings.cpp:166-179) on the context's clock; host-side helper. */
double vrt_update_time(double time, int step, double dt);
/* SolverManager::Advance(dt) (SolverManager.cpp:28-39): the six stages, captured once as a CUDA graph and
 * replayed.  laser[2*i], laser[2*i+1] = GetBY, GetBZ(0, time after UpdateTime(i)).  Advances VRT_TIME. */
int vrt_step(vrt_ctx* ctx, double dt, const double laser[12]);
/* SolverManager::AdvanceFields(dt) (SolverManager.cpp:41-46). */
int vrt_step_fields(vrt_ctx* ctx, double dt, const double laser[12]);
/* number of kernels the last vrt_step / vrt_step_fields launched (for bench.py's gpu_launches) */
long vrt_last_step_launches(const vrt_ctx* ctx);

/* Diagnostic (no reference counterpart): the launch plan the fused streaming path uses for species s, so that a test can assert
 * which kernel specialisations a mesh exercises.  out[0] = CTA width W (threads; 128 = the compile-time-width instance),
 * out[1] = p strips, out[2] = x chunks, out[3] = CTAs taking the interior specialisation (no boundary predicates),
 * out[4], out[5] = cells per thread and threads per CTA of the streaming moments kernel (k_slab_moments<CPT, NT, .>). */
int vrt_fused_plan(vrt_ctx* ctx, int s, int out[6]);

/* option 0: replay vrt_step through a captured CUDA graph (default 1) or launch kernel by kernel (0);
 * option 1: run the species' Vlasov stages of one RK stage on concurrent streams / graph branches (default 1);
 * option 2: run the 1-D Maxwell stage of an RK stage as a branch concurrent with that stage's Poisson solve and Vlasov kernels
 *           (default 1; it depends on the moments only, and writes a^2 into a second buffer that is exchanged at the join) */
int vrt_set_option(vrt_ctx* ctx, int option, int value);
/* Rectangle::InitializeDistribution (Rectangle.cpp:616-665) for the shipped Maxwellian slab
 * (Settings::InitialDistribution, veritas.cpp:107-115), evaluated on the device: sub-cell midpoint quadrature
 * with r^(depth+quadrature_depth) points per direction; writes states 0 and 1.  SURVEY.md §8(f) item 2. */
int vrt_init_maxwellian_slab(vrt_ctx* ctx, int s, double xl, double xr, double n0, double T, int quadrature_depth);

/* Rectangle::CalculateEnergy (Rectangle.cpp:284-305), the energy-spectrum diagnostic dN/dp: energyR of one patch (n_p * r^depth
 * values on the finest p grid: dx * sum over the patch's non-nested cells of the p sub-cell interpolants of state 1).  The
 * reference accumulates it with a data race; this sum is deterministic.  Level::CollectEnergy,
 * Mesh::InterpolateEnergyToFinestMesh and EMFieldSolver::AssembleEnergy / DumpEnergy stay on the host (tiny arrays). */
int vrt_patch_energy(vrt_ctx* ctx, int s, int patch, double* energy_host);

/* ---- checkpoint / restart (no counterpart in the reference) ---------------------------- */
/* Binary image of the context at a step boundary: hierarchy and f of every species, the 1-D field arrays, PHI, Ex0, the
 * neutralisation charge and the time.  vrt_checkpoint_read needs a context with the same vrt_set_grid (and vrt_set_slab /
 * vrt_set_path) calls; it recreates the hierarchies itself.  A restarted run continues bit for bit.  One file per rank. */
/* The hierarchy species s holds — the descriptors of the last vrt_set_hierarchy / vrt_regrid / vrt_checkpoint_read, in the caller's
 * order: returns the patch count (negative: error) and fills at most `capacity` entries of `out` (NULL with capacity 0 to query the
 * count).  Lets a host layer rebuild its Level / Rectangle objects after vrt_checkpoint_read (SolverManager::Restart). */
int vrt_get_hierarchy(vrt_ctx* ctx, int s, int capacity, vrt_patch_desc* out);
int vrt_checkpoint_write(vrt_ctx* ctx, const char* path);
int vrt_checkpoint_read(vrt_ctx* ctx, const char* path);

/* ---- hierarchy connectivity on the host (no device needed) ---------------------------------------- */
/* The tables Rectangle::CalculateConnectivitySame / CalculateConnectivityFromFiner build (Rectangle.cpp:671-864), as
 * vrt_set_hierarchy derives them.  side: 0 xm, 1 xp (n_p/r + 2 strips, first and last = the 2-cell corners), 2 pm, 3 pp
 * (n_x/r strips).  nb = neighbour patch number in the caller's numbering, -1 = the BoundaryCondition object;
 * same = is_exterrior_boundary_same_level_*.  flags per padded cell: 1 is_nested, 2 / 4 is_interrior_level_boundary_x / _p. */
typedef struct vrt_conn vrt_conn;
int vrt_conn_create(vrt_conn** out, int n_patches, const vrt_patch_desc* patches, int refinement_ratio, int max_depth);
int vrt_conn_strips(const vrt_conn* conn, int patch, int side, int* nb, unsigned char* same);   /* returns the strip count */
int vrt_conn_flags(const vrt_conn* conn, int patch, unsigned char* flags_padded);
void vrt_conn_destroy(vrt_conn* conn);

/* ---- multi-GPU x-slabs (no counterpart in the reference) ------------------- */
/* This context owns finest columns [x_begin, x_end) of every full-domain single-level patch. */
int vrt_set_slab(vrt_ctx* ctx, int rank, int n_ranks, int x_begin, int x_end);
int vrt_nccl_unique_id(void* out128);
int vrt_comm_init(vrt_ctx* ctx, const void* unique_id128, int rank, int n_ranks);

/* ---- laser-plasma case helpers (veritas.cpp:36-115 restated in host/laser_plasma_case.hpp) --- */
typedef struct {
    double lambda, a0, density, temp_frac, pmax_e, pmax_i, box_lambdas, ion_mass_ratio;
} vrt_case_params;
typedef struct {
    double dp[2], pmin[2], dx, sizeWeight, temp0[2], temp1[2], tempEM[2];
    int quadratureDepth;
} vrt_case_derived;
int vrt_case_derive(const vrt_case_params* p, double m0, double q0, unsigned x_size, const unsigned* p_size,
                    double refinementCriteria, vrt_case_derived* out);
double vrt_case_laser_by(double lambda, double amp, double x, double t);
double vrt_case_laser_bz(double lambda, double amp, double x, double t);
double vrt_case_maxwellian_slab(double x, double p, double xl, double xr, double n0, double T);

#ifdef __cplusplus
}
#endif
#endif /* VERITAS_B200_H */
