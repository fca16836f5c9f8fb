// veritas_b200 internal declarations shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/veritas_b200.h"

#define VRT_EPS0_INV 1.1294e+11     // veritas.hpp:22 (reference literals, quirk Q9)
#define VRT_MU_INV 795774.715482    // veritas.hpp:25
#define VRT_CS 299792458.0          // veritas.hpp:26
#define VRT_C_INV 3.33564095e-9     // veritas.hpp:27
#define VRT_SLAB_GH 5               // ghost doubles below p-cell 0 in a slab column (odd: see vrt_set_hierarchy)

// RK tableau, literal values of Rectangle.cpp:1397-1498 / EMSolver.cpp:210-312 (row s = stage s, row 5 = b)
struct VrtTableau { double a[6][6]; };
static const VrtTableau kTableau = {{{0.5, 0, 0, 0, 0, 0},
                                     {0.221776, 0.110224, 0, 0, 0, 0},
                                     {-0.04884659515311857, -0.17772065232640102, 0.8465672474795197, 0, 0, 0},
                                     {-0.15541685842491548, -0.3567050098221991, 1.0587258798684427, 0.30339598837867193, 0, 0},
                                     {0.2014243506726763, 0.008742057842904185, 0.15993995707168115, 0.4038290605220775, 0.22606457389066084, 0},
                                     {0.15791629516167136, 0.0, 0.18675894052400077, 0.6805652953093346, -0.27524053099500667, 0.25}}};

// Per-step parameters living in device memory so that a captured CUDA graph can be replayed with new values:
// written by one cudaMemcpyAsync from pinned host memory before the graph launch.
struct VrtStepParams {
    double dt;
    double laser[12];   // by0,bz0 per stage
};

// 1-D field-solver state on the device (EMFieldSolver members, EMSolver.hpp:10-23)
struct VrtFields {
    int N, pre, post, M;        // x_size_finest, n_prepad, n_postpad, M = N+pre+post
    double dx;                  // finest dx
    double* Y[6];               // By,Bz,Ey,Ez,Ay,Az: 8 slots x M, slot-major (Index(i,s) = s*M+i)
    double* a_squared;          // N+1 (x-faces; entry N never written, quirk Q2)
    double* PHI;                // N
    int epad;                   // E is tabulated on [-epad, N+epad): epad = max(2, r^max_depth) (coarse ghost columns average rtb cells)
    double* E;                  // N+2*epad: E[i+epad] = EMFieldSolver::GetEfield(i)
    double* charge; double* J; double* neutral;   // N each
    double* Ex0;                // device scalar
    double* scratch;            // Poisson workspace, 4*N
    double* cfl;                // device scalar (max-reduction result)
};

// what the assembly kernel needs to know about one species' moments
struct VrtAssembleSpecies {
    int mode;                          // 0: slab arrays, 1: patch table
    const double *chargeR, *currentR;  // mode 0: the slab's columns [x0, x0 + n)
    int x0, n;
    const struct VrtPatchDev* all;     // mode 1: device patch table and its level ranges (finest first)
    int n_levels, first[16], count[16];
    double* charges;                   // out: N entries
};
struct VrtAssembleArgs { int n_species, N, total; VrtAssembleSpecies sp[8]; double* J; double* charge; };

// Species constants (Settings.hpp:39-41)
struct VrtSpecies {
    double m, q, pmin, dp_finest;
};

// One Rectangle on the device, split path: SoA planes in the reference's padded layout.
struct VrtPatchDev {
    int n_x, n_p, x_pos, p_pos, up, down, left, right, rtb, depth;
    int pitch;                   // n_p + 4
    long npad;                   // (n_x+4)*(n_p+4)
    double dx, dp;
    double *f0, *f1, *f2, *fx, *fp, *ex, *ep;
    double *FxH, *FpH;           // 6 planes each, slot-major
    double *FxL, *FpL;           // slot 0 (quirk Q1)
    double *FxLS, *FpLS, *FxDS, *FpDS, *Rp, *Rm, *Cx, *Cp;
    double *chargeR, *currentR;  // n_x*rtb each
    // AMR connectivity (vrt_amr.cu), derived from the patch descriptors as Rectangle::CalculateConnectivitySame /
    // CalculateConnectivityFromFiner do (Rectangle.cpp:671-864).  Neighbour indices are device-table indices, -1 = the
    // BoundaryCondition object.  Side 0 xm, 1 xp (n_p/r + 2 strips incl. the two corners), 2 pm, 3 pp (n_x/r strips).
    int ns_x, ns_p;
    int* nb[4]; unsigned char* same[4];
    int *finer, *finer_x, *finer_p;   // per padded cell: finer patch covering the cell / owning the flagged x- or p-face
    unsigned char* flags;             // VRT_NESTED | VRT_LBX | VRT_LBP per padded cell
};
enum { VRT_NESTED = 1, VRT_LBX = 2, VRT_LBP = 4 };
// one coarse face whose flux is replaced by the finer patch's (is_interrior_level_boundary_{x,p}): table index of the coarse patch,
// direction (0: x-face, 1: p-face), padded cell index — the list lets the flux-matching kernel launch exactly the work there is
struct VrtLbFace { int patch; int dir; long cell; };

// Host-side connectivity tables of one species' hierarchy (caller's patch numbering unless noted)
struct VrtConnPatch {
    int ns_x = 0, ns_p = 0;
    std::vector<int> nb[4];
    std::vector<unsigned char> same[4];
    std::vector<int> finer, finer_x, finer_p;
    std::vector<unsigned char> flags;
};
struct vrt_conn {
    int r = 2, max_depth = 0;
    std::vector<vrt_patch_desc> desc;
    std::vector<VrtConnPatch> P;
    std::string err;
};
int vrt_conn_derive(vrt_conn& C, int n, const vrt_patch_desc* d, int r, int max_depth);

// Fused-path storage of one full-domain (or x-slab) single-level patch: three rotating f planes, five
// stored high-order flux pairs and the low-order pair of stage 0.  Rows are x columns (slow), p is contiguous.  GX ghost columns per side.
struct VrtSlabDev {
    int n_x, n_p;                // local interior columns, p cells
    int x_begin;                 // global finest index of local column 0
    int n_x_global;
    int left, right;             // this slab touches the physical x boundary
    int gx;                      // ghost columns per side (3)
    int pitch;                   // doubles per column (n_p + 4 rounded up)
    long plane;                  // doubles per plane = (n_x + 2*gx) * pitch
    double dx, dp;
    double* f[3];                // rotating: cur0 = f^n, cur1 = stage value, spare
    double* FxH[5]; double* FpH[5];   // one allocation, planes interleaved FxH[0], FpH[0], FxH[1], ... (one 3-D TMA box covers a stage's history)
    double *FxL0, *FpL0;         // planes 10, 11 of the same allocation: the unscaled low-order flux pair of stage 0 (quirk Q1)
    double *chargeR, *currentR;  // n_x each
};
// TMA descriptors (CUtensorMap, 128 bytes each) of a slab's planes for the fused stage: [0] the three f planes with a
// one-column box, [S] (S = 1..5) the interleaved flux history with a box of 2S planes
struct alignas(64) VrtSlabMaps { unsigned char m[6][128]; int W; };

struct VrtSpeciesState {
    VrtSpecies sp;
    bool configured = false;
    int path = VRT_PATH_SPLIT;
    std::vector<vrt_patch_desc> desc;
    // split path
    std::vector<VrtPatchDev> patches;    // host copy of the descriptors (device pointers inside), caller's numbering
    std::vector<VrtPatchDev> table;      // the same, grouped by depth = order of the device table
    std::vector<int> table_index;        // caller's patch number -> table index
    std::vector<int> table_order;        // table index -> caller's patch number
    VrtPatchDev* d_patches = nullptr;    // device copy of `table`
    std::vector<double*> allocations;
    VrtSlabMaps maps;
    void* conn_pool = nullptr;           // device pool holding the connectivity tables of all patches
    VrtLbFace* d_lb = nullptr;           // (inside conn_pool) the flagged coarse faces of all levels, grouped by depth like the table
    std::vector<int> lb_first, lb_count; // per depth: range in d_lb
    bool has_amr = false;                // any nested cell / coarse-fine face / same-level neighbour
    std::vector<std::vector<int>> level_patches;   // table indices per depth (contiguous ranges)
    // fused path
    VrtSlabDev slab;
    int i_f0 = 0, i_f1 = 0;              // indices into slab.f: f^n and current stage value
    double* d_charges = nullptr;         // per-species charge on the finest grid (N)
    // x-slab runs: the halo exchange of this species runs on the context's communication stream behind ev_k (stage kernel done)
    // and signals ev_h; the next stage kernel of the species waits for it (halo_pending)
    cudaEvent_t ev_k = nullptr, ev_h = nullptr;
    bool halo_pending = false;
};

struct vrt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int n_species = 0;
    int refinement_ratio = 2, max_depth = 0;
    bool grid_set = false;
    VrtFields F{};
    std::vector<double*> field_allocs;
    std::vector<VrtSpeciesState> S;
    int requested_path = VRT_PATH_AUTO;
    double time = 0.0;
    // slab decomposition
    int rank = 0, n_ranks = 1, x_begin = 0, x_end = 0;
    void* nccl_comm = nullptr;
    // every NCCL call of the context is issued on comm_stream, in the same order on all ranks (halo s = 0, 1, ..., moment
    // all-gather, ...), so that the exchanges overlap the other species' stage kernel, the 1-D field update and the moments
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_m = nullptr, ev_ag = nullptr;
    // step graph
    VrtStepParams* d_params = nullptr; VrtStepParams* h_params = nullptr;
    cudaGraphExec_t graph_step3[3] = {nullptr, nullptr, nullptr};   // keyed by the plane-rotation state at step start
    // executables of the step graph a regrid invalidated: the re-captured graph has the same topology as long as the same levels are
    // populated, so the old executable is UPDATED with the new nodes' parameters (cudaGraphExecUpdate) instead of instantiating a new
    // one — instantiation is the larger part of what a regrid costs a small hierarchy
    cudaGraphExec_t graph_stale[3] = {nullptr, nullptr, nullptr};
    cudaGraphExec_t graph_fields = nullptr;
    long graph_launches[3] = {0, 0, 0}, graph_fields_launches = 0;
    std::pair<int, int> graph_end_state[3][8];
    bool use_graph = true;
    long launches = 0, last_step_launches = 0;
    double* d_comm = nullptr; long comm_doubles = 0;   // staging buffer for the moment all-gather
    // the species' Vlasov stages of one RK stage are independent: species s > 0 runs on aux_stream[s - 1] between a fork and a
    // join event (parallel branches of the step graph; on x-slab runs the NCCL calls still go to comm_stream in host order)
    std::vector<cudaStream_t> aux_stream;
    std::vector<cudaEvent_t> aux_join;
    cudaEvent_t ev_fork = nullptr;
    bool fork_species = true;
    // the 1-D Maxwell stage of an RK stage depends on the moments only (J) and is independent of that stage's Poisson solve and
    // Vlasov kernels: inside vrt_step it runs on field_stream between ev_ffork and ev_fjoin and writes a^2 at the faces into
    // asq_alt while the Vlasov kernels still read F.a_squared; the two pointers are exchanged at the join (six times per step)
    cudaStream_t field_stream = nullptr;
    cudaEvent_t ev_ffork = nullptr, ev_fjoin = nullptr;
    double* asq_alt = nullptr;
    bool fork_fields = true;
};

#define VRT_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return VRT_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

// ---- helpers of vrt_abi.cu used by other units (internal; C linkage only because they are defined inside its extern "C" block)
extern "C" {
int vrt_hierarchy_path(vrt_ctx* c, int n_patches, const vrt_patch_desc* d, int* path_out);
long vrt_slab_plane_doubles(int n_x_local, int n_p);
void vrt_invalidate_hierarchies(vrt_ctx* c);
}

// ---- launchers implemented in the kernel translation units -----------------------------------------
// split path (vrt_split.cu, compiled with -fmad=false)
int vrt_split_substep(vrt_ctx* c, int s, int depth, const double* d_dt, int step, int substep);
int vrt_split_substep_all(vrt_ctx* c, int s, const double* d_dt, int step, int substep);
int vrt_split_moments(vrt_ctx* c, int s);      // the species' per-patch kernel only (chargeR / currentR of every patch)
int vrt_split_patch_energy(vrt_ctx* c, int s, int patch, double* host_energy);
// AMR kernels (vrt_amr.cu, compiled with -fmad=false)
int vrt_amr_upload_connectivity(vrt_ctx* c, int s, const vrt_conn& C);
int vrt_amr_level_pass(vrt_ctx* c, int s, int depth, int type, int val);
int vrt_amr_push_data(vrt_ctx* c, int s, int val);
int vrt_amr_push_boundary_c(vrt_ctx* c, int s);
int vrt_amr_level_boundary_fluxes(vrt_ctx* c, int s, int depth, int step);
int vrt_amr_transfer(vrt_ctx* c, const VrtSpeciesState& old_state, VrtSpeciesState& new_state);
int vrt_amr_error_flags(vrt_ctx* c, int s, int patch, const double weights[5], double criteria, unsigned char* flags_host);
// 1-D solver (vrt_fields.cu, compiled with -fmad=false)
int vrt_fields_rhs_update_faces(vrt_ctx* c, int step, const VrtStepParams* d_params, double* asq_out = nullptr);
int vrt_fields_poisson(vrt_ctx* c);
int vrt_fields_cfl(vrt_ctx* c);
// EMFieldSolver::AssembleRhoAndJ (EMSolver.cpp:104-122) behind the species' moment kernels: ONE launch forms, per finest column,
// charges[s] of every species in `mask` (bit s), the current J accumulated over those species in species order, and (total != 0)
// charge = sum of the species charges — each from zero, with the additions and their order of the reference's loops
int vrt_fields_assemble(vrt_ctx* c, unsigned mask, int total);
int vrt_fields_total_charge(vrt_ctx* c);      // charge = sum_s charges[s] (x-slab runs: after the all-gather)
int vrt_fields_neutralize(vrt_ctx* c);
// fused path (vrt_fused.cu)
int vrt_fused_stage(vrt_ctx* c, int s, const double* d_dt, int step);
int vrt_fused_moments(vrt_ctx* c, int s);      // the species' slab kernel only (chargeR / currentR of the slab's columns)
int vrt_fused_zero_ghosts(vrt_ctx* c, int s, int plane_idx);
int vrt_fused_make_maps(vrt_ctx* c, int s);
int vrt_fused_plan_impl(vrt_ctx* c, int s, int out[6]);
