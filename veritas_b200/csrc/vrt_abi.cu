// C ABI of veritas_b200 (include/veritas_b200.h): context, device storage, orchestration of the stage loop.
#include "vrt_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <stdexcept>

int vrt_fields_refresh_efield(vrt_ctx* c);
int vrt_init_kernels_maxwellian(vrt_ctx* c, int s, double xl, double xr, double n0, double T, int quadrature_depth);
int vrt_comm_halo_exchange(vrt_ctx* c, int s);
int vrt_comm_gather_moments(vrt_ctx* c);
int vrt_comm_wait_halo(vrt_ctx* c, int s);
void vrt_comm_destroy(vrt_ctx* c);
int vrt_fields_init_tables(vrt_ctx* c);
int vrt_split_init_tables(vrt_ctx* c);

static std::string g_err;
extern "C" { static int ready_species(vrt_ctx* c, int s); }

namespace {

__global__ void k_set_params(VrtStepParams* p, VrtStepParams v) { *p = v; }
__global__ void k_fill(double* a, long n, double v) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

int dev_alloc(vrt_ctx* c, std::vector<double*>& pool, double** out, size_t n_doubles) {
    double* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n_doubles, 1) * sizeof(double));
    if (e != cudaSuccess) { c->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return VRT_ERR_NOMEM; }
    e = cudaMemsetAsync(p, 0, std::max<size_t>(n_doubles, 1) * sizeof(double), c->stream);
    if (e != cudaSuccess) { c->err = std::string("cudaMemset: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
    pool.push_back(p);
    *out = p;
    return 0;
}

void free_species(VrtSpeciesState& S) {
    for (double* p : S.allocations) cudaFree(p);
    S.allocations.clear();
    if (S.d_patches) { cudaFree(S.d_patches); S.d_patches = nullptr; }
    if (S.conn_pool) { cudaFree(S.conn_pool); S.conn_pool = nullptr; }
    S.d_lb = nullptr; S.lb_first.clear(); S.lb_count.clear();
    S.has_amr = false;
    S.patches.clear(); S.level_patches.clear(); S.desc.clear();
    S.configured = false;
}

// keep_for_update: the step executables are parked in graph_stale (see vrt_internal.cuh) instead of being destroyed
void drop_graphs(vrt_ctx* c, bool keep_for_update = false) {
    for (int k = 0; k < 3; k++) {
        if (!c->graph_step3[k]) continue;
        if (keep_for_update && !c->graph_stale[k]) c->graph_stale[k] = c->graph_step3[k];
        else cudaGraphExecDestroy(c->graph_step3[k]);
        c->graph_step3[k] = nullptr;
    }
    if (!keep_for_update) for (int k = 0; k < 3; k++) if (c->graph_stale[k]) { cudaGraphExecDestroy(c->graph_stale[k]); c->graph_stale[k] = nullptr; }
    if (c->graph_fields) { cudaGraphExecDestroy(c->graph_fields); c->graph_fields = nullptr; }
}

bool check(vrt_ctx* c, bool ok, const char* msg) { if (!ok) c->err = msg; return ok; }

int set_params_async(vrt_ctx* c, double dt, const double* laser) {
    VrtStepParams v{}; v.dt = dt;
    if (laser) std::memcpy(v.laser, laser, sizeof(v.laser));
    k_set_params<<<1, 1, 0, c->stream>>>(c->d_params, v);
    VRT_CUDA(c, cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char* vrt_global_error(void) { return g_err.c_str(); }
const char* vrt_last_error(const vrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
const char* vrt_version(void) { return "veritas_b200 0.1 (sm_100a)"; }

int vrt_create(vrt_ctx** out, int device, int n_species) {
    if (!out || n_species < 1 || n_species > 8) { g_err = "vrt_create: bad arguments"; return VRT_ERR_ARG; }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_err = std::string("vrt_create: no CUDA device (") + cudaGetErrorString(e) + "); veritas_b200 has no CPU fallback";
        return VRT_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) { g_err = "vrt_create: device ordinal out of range"; return VRT_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
    vrt_ctx* c = new vrt_ctx();
    c->device = device; c->n_species = n_species; c->S.resize(n_species);
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_params, sizeof(VrtStepParams));
    if (e == cudaSuccess) e = cudaMemset(c->d_params, 0, sizeof(VrtStepParams));
    if (e != cudaSuccess) { g_err = std::string("vrt_create: ") + cudaGetErrorString(e); delete c; return VRT_ERR_CUDA; }
    if (vrt_fields_init_tables(c) || vrt_split_init_tables(c)) { g_err = "vrt_create: " + c->err; delete c; return VRT_ERR_CUDA; }
    *out = c;
    return 0;
}

int vrt_destroy(vrt_ctx* c) {
    if (!c) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    vrt_comm_destroy(c);
    for (auto& S : c->S) { free_species(S); if (S.d_charges) cudaFree(S.d_charges); }
    for (double* p : c->field_allocs) cudaFree(p);
    if (c->d_params) cudaFree(c->d_params);
    if (c->d_comm) cudaFree(c->d_comm);
    if (c->field_stream) cudaStreamDestroy(c->field_stream);
    if (c->ev_ffork) cudaEventDestroy(c->ev_ffork);
    if (c->ev_fjoin) cudaEventDestroy(c->ev_fjoin);
    for (cudaStream_t st : c->aux_stream) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t ev : c->aux_join) if (ev) cudaEventDestroy(ev);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int vrt_sync(vrt_ctx* c) {
    if (!c) return VRT_ERR_ARG;
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
void* vrt_stream(vrt_ctx* c) { return c ? (void*)c->stream : nullptr; }

int vrt_set_grid(vrt_ctx* c, int N, double dx, int pre, int post, int r, int max_depth) {
    if (!c) return VRT_ERR_ARG;
    if (!check(c, N >= 8 && dx > 0 && pre >= 2 && post >= 2 && r >= 2 && max_depth >= 0, "vrt_set_grid: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, !c->grid_set, "vrt_set_grid: grid already set")) return VRT_ERR_STATE;
    cudaSetDevice(c->device);
    VrtFields& F = c->F;
    F.N = N; F.pre = pre; F.post = post; F.M = N + pre + post; F.dx = dx;
    c->refinement_ratio = r; c->max_depth = max_depth;
    // the coarse-fine flux matching recurses once per level (vrt_amr.cu: rgk_flux, 280-byte frames)
    if (max_depth >= 2) cudaDeviceSetLimit(cudaLimitStackSize, 1024 + 384 * (size_t)max_depth);
    int rc;
    for (int v = 0; v < 6; v++) if ((rc = dev_alloc(c, c->field_allocs, &F.Y[v], 8L * F.M))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.a_squared, N + 1))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &c->asq_alt, N + 1))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.PHI, N))) return rc;
    F.epad = std::max(2, (int)std::lround(std::pow((double)r, max_depth)));
    if (!check(c, F.epad < N / 2, "vrt_set_grid: too many levels for this x size")) return VRT_ERR_ARG;
    if ((rc = dev_alloc(c, c->field_allocs, &F.E, N + 2 * F.epad))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.charge, N))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.J, N))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.neutral, N))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.Ex0, 2))) return rc;
    if ((rc = dev_alloc(c, c->field_allocs, &F.scratch, 4L * N + 8 + 4 * 4096))) return rc;   // 3 N-vectors + sum(b) + Poisson tile partials (vrt_fields.cu); level moments use the first 2N
    if ((rc = dev_alloc(c, c->field_allocs, &F.cfl, 2))) return rc;
    for (auto& S : c->S) { double* p; if ((rc = dev_alloc(c, c->field_allocs, &p, N))) return rc; S.d_charges = p; c->field_allocs.pop_back(); }
    c->x_begin = 0; c->x_end = N;
    c->grid_set = true;
    return 0;
}

int vrt_set_species(vrt_ctx* c, int s, double mass, double charge, double pmin, double dp_finest) {
    if (!c) return VRT_ERR_ARG;
    if (!check(c, s >= 0 && s < c->n_species && mass > 0 && dp_finest > 0, "vrt_set_species: bad arguments")) return VRT_ERR_ARG;
    c->S[s].sp = VrtSpecies{mass, charge, pmin, dp_finest};
    c->S[s].configured = true;
    return 0;
}

int vrt_set_path(vrt_ctx* c, int path) {
    if (!c || path < VRT_PATH_AUTO || path > VRT_PATH_FUSED) return VRT_ERR_ARG;
    c->requested_path = path;
    return 0;
}
int vrt_get_path(vrt_ctx* c, int s) { return (c && s >= 0 && s < c->n_species) ? c->S[s].path : VRT_ERR_ARG; }

int vrt_set_slab(vrt_ctx* c, int rank, int n_ranks, int x_begin, int x_end) {
    if (!c) return VRT_ERR_ARG;
    if (!check(c, c->grid_set, "vrt_set_slab: call vrt_set_grid first")) return VRT_ERR_STATE;
    if (!check(c, n_ranks >= 1 && rank >= 0 && rank < n_ranks && x_begin >= 0 && x_end <= c->F.N && x_end - x_begin >= 8,
               "vrt_set_slab: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, n_ranks == 1 || ((x_end - x_begin) * n_ranks == c->F.N && x_begin == rank * (x_end - x_begin)),
               "vrt_set_slab: slabs must be equal and ordered by rank")) return VRT_ERR_ARG;
    for (auto& S : c->S) if (!check(c, S.desc.empty(), "vrt_set_slab: must precede vrt_set_hierarchy")) return VRT_ERR_STATE;
    c->rank = rank; c->n_ranks = n_ranks; c->x_begin = x_begin; c->x_end = x_end;
    return 0;
}

// split path storage of one species: SoA planes per patch in the reference's padded layout, the device patch table grouped
// by depth, and the connectivity tables (S must hold no storage; the caller sets S.desc once this has succeeded)
static int build_split(vrt_ctx* c, int s, int n_patches, const vrt_patch_desc* d) {
    VrtSpeciesState& S = c->S[s];
    const VrtSpecies sp = S.sp;
    const int r = c->refinement_ratio;
    int rc;
    S.level_patches.assign(c->max_depth + 1, {});
    std::vector<int> order(n_patches);
    for (int p = 0; p < n_patches; p++) order[p] = p;
    // the device table is grouped by depth so that one level is a contiguous range; S.patches is indexed by
    // the caller's patch number, table_index maps it into the table
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a].depth < d[b].depth; });
    vrt_conn conn;
    if (vrt_conn_derive(conn, n_patches, d, r, c->max_depth)) { c->err = "vrt_set_hierarchy: " + conn.err; return VRT_ERR_ARG; }
    S.patches.resize(n_patches);
    std::vector<VrtPatchDev> table(n_patches);
    S.table_index.assign(n_patches, 0);
    S.table_order = order;
    // ONE pooled, zeroed allocation for the planes of every patch (17 single planes, two 6-slot flux histories, the two 1-D moment
    // arrays; each plane 256-byte aligned): a regrid then costs one cudaMalloc per species instead of 21 per patch
    auto aligned = [](size_t n_doubles) { return (n_doubles + 31) & ~(size_t)31; };
    size_t total = 0;
    for (int p = 0; p < n_patches; p++) {
        const vrt_patch_desc& q = d[p];
        const size_t npad = (size_t)(q.n_x + 4) * (q.n_p + 4);
        const size_t rtb = (size_t)std::lround(std::pow((double)r, q.depth));
        total += 17 * aligned(npad) + 2 * aligned(6 * npad) + 2 * aligned((size_t)q.n_x * rtb);
    }
    double* pool = nullptr;
    if ((rc = dev_alloc(c, S.allocations, &pool, total))) return rc;
    size_t used = 0;
    auto take = [&](double** out, size_t n_doubles) { *out = pool + used; used += aligned(n_doubles); };
    for (int ti = 0; ti < n_patches; ti++) {
        const int p = order[ti];
        const vrt_patch_desc& q = d[p];
        VrtPatchDev P{};
        P.n_x = q.n_x; P.n_p = q.n_p; P.x_pos = q.x_pos; P.p_pos = q.p_pos;
        P.up = q.up; P.down = q.down; P.left = q.left; P.right = q.right; P.depth = q.depth;
        P.rtb = (int)std::lround(std::pow((double)r, q.depth));
        P.pitch = q.n_p + 4; P.npad = (long)(q.n_x + 4) * (q.n_p + 4);
        P.dx = std::pow((double)r, (double)q.depth) * c->F.dx;              // Settings::GetDx (Settings.cpp:142-144)
        P.dp = std::pow((double)r, (double)q.depth) * sp.dp_finest;         // Settings::GetDp (Settings.cpp:138-140)
        double** planes[] = {&P.f0, &P.f1, &P.f2, &P.fx, &P.fp, &P.ex, &P.ep, &P.FxL, &P.FpL, &P.FxLS, &P.FpLS, &P.FxDS, &P.FpDS, &P.Rp, &P.Rm, &P.Cx, &P.Cp};
        for (double** pl : planes) take(pl, P.npad);
        take(&P.FxH, 6 * P.npad);
        take(&P.FpH, 6 * P.npad);
        take(&P.chargeR, (size_t)P.n_x * P.rtb);
        take(&P.currentR, (size_t)P.n_x * P.rtb);
        S.patches[p] = P; table[ti] = P; S.table_index[p] = ti;
        S.level_patches[q.depth].push_back(ti);
    }
    S.table = table;
    if ((rc = vrt_amr_upload_connectivity(c, s, conn))) return rc;
    VRT_CUDA(c, cudaMalloc(&S.d_patches, sizeof(VrtPatchDev) * n_patches));
    VRT_CUDA(c, cudaMemcpyAsync(S.d_patches, S.table.data(), sizeof(VrtPatchDev) * n_patches, cudaMemcpyHostToDevice, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}


// Validation of a hierarchy and the path it would take, without touching any state (vrt_set_hierarchy and the first pass of
// vrt_checkpoint_read, which must not modify the context before the whole file has been checked).
int vrt_hierarchy_path(vrt_ctx* c, int n_patches, const vrt_patch_desc* d, int* path_out) {
    const int r = c->refinement_ratio;
    for (int p = 0; p < n_patches; p++) {
        const vrt_patch_desc& q = d[p];
        if (!check(c, q.depth >= 0 && q.depth <= c->max_depth && q.n_x >= r && q.n_p >= r && q.n_x % r == 0 && q.n_p % r == 0,
                   "vrt_set_hierarchy: bad patch descriptor")) return VRT_ERR_ARG;
    }
    // path selection: the fused streaming kernel serves single-level full-domain patches (optionally x-slabs)
    const vrt_patch_desc& q0 = d[0];
    const bool full = (n_patches == 1 && c->max_depth == 0 && q0.depth == 0 && q0.x_pos == 0 && q0.p_pos == 0 && q0.n_x == c->F.N &&
                       q0.up && q0.down && q0.left && q0.right);
    int path = c->requested_path;
    if (path == VRT_PATH_AUTO) path = full ? VRT_PATH_FUSED : VRT_PATH_SPLIT;
    if (!check(c, path != VRT_PATH_FUSED || full, "vrt_set_hierarchy: the fused path needs one full-domain single-level patch")) return VRT_ERR_ARG;
    if (!check(c, c->n_ranks == 1 || path == VRT_PATH_FUSED, "vrt_set_hierarchy: x-slabs need the fused path")) return VRT_ERR_ARG;
    *path_out = path;
    return 0;
}
// doubles per slab plane of the fused layout for n_x local columns of n_p cells (vrt_set_hierarchy; vrt_checkpoint_read sizes its
// records with it): even pitch, 3 ghost columns per side, slack for the last strip's bulk load
long vrt_slab_plane_doubles(int n_x_local, int n_p) {
    const long pitch = ((n_p + VRT_SLAB_GH + 4 + 1) / 2) * 2;
    return (long)(n_x_local + 2 * 3) * pitch + 1024;
}
// a context whose hierarchies could not be (re)built holds none: every later call fails with VRT_ERR_STATE instead of touching
// half-built storage
void vrt_invalidate_hierarchies(vrt_ctx* c) {
    drop_graphs(c);
    for (auto& S : c->S) { VrtSpecies sp = S.sp; const bool conf = S.configured; free_species(S); S.sp = sp; S.configured = conf; }
}

static int set_hierarchy_impl(vrt_ctx* c, int s, int n_patches, const vrt_patch_desc* d);

int vrt_set_hierarchy(vrt_ctx* c, int s, int n_patches, const vrt_patch_desc* d) {
    if (!c) return VRT_ERR_ARG;
    if (!check(c, c->grid_set && s >= 0 && s < c->n_species && c->S[s].configured, "vrt_set_hierarchy: set grid and species first")) return VRT_ERR_STATE;
    if (!check(c, n_patches >= 1 && d, "vrt_set_hierarchy: bad arguments")) return VRT_ERR_ARG;
    int rc;
    try { rc = set_hierarchy_impl(c, s, n_patches, d); }
    catch (const std::exception&) { c->err = "vrt_set_hierarchy: out of host memory"; rc = VRT_ERR_NOMEM; }
    if (rc) {      // no half-built species: storage released, descriptor list empty (ready() then answers VRT_ERR_STATE)
        VrtSpeciesState& S = c->S[s];
        VrtSpecies sp = S.sp;
        const std::string msg = c->err;
        free_species(S);
        S.sp = sp; S.configured = true;
        c->err = msg;
    }
    return rc;
}

static int set_hierarchy_impl(vrt_ctx* c, int s, int n_patches, const vrt_patch_desc* d) {
    cudaSetDevice(c->device);
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    VrtSpeciesState& S = c->S[s];
    VrtSpecies sp = S.sp;
    free_species(S);
    S.sp = sp; S.configured = true;
    int path;
    if (int rc = vrt_hierarchy_path(c, n_patches, d, &path)) return rc;
    const vrt_patch_desc& q0 = d[0];
    S.path = path;
    int rc;
    if (path == VRT_PATH_FUSED) {
        VrtSlabDev& L = S.slab;
        L = VrtSlabDev{};
        L.n_x = c->x_end - c->x_begin; L.n_p = q0.n_p; L.x_begin = c->x_begin; L.n_x_global = c->F.N;
        L.left = (c->x_begin == 0); L.right = (c->x_end == c->F.N);
        // column layout: VRT_SLAB_GH ghost doubles, n_p cells, >= 4 ghost doubles; even pitch keeps every strip start
        // (VRT_SLAB_GH + j0 - 3, j0 even) 16-byte aligned for the bulk-async (TMA) column loads
        L.gx = 3; L.pitch = ((q0.n_p + VRT_SLAB_GH + 4 + 1) / 2) * 2;
        L.plane = vrt_slab_plane_doubles(L.n_x, q0.n_p);       // incl. slack: the last strip's bulk load may run past the last column
        if (!check(c, L.plane < (1L << 31), "vrt_set_hierarchy: slab plane exceeds 2^31 cells (split the domain over more GPUs)")) return VRT_ERR_ARG;
        L.dx = c->F.dx; L.dp = sp.dp_finest;
        // two pooled allocations (uniform plane stride): each is one 3-D tensor {p, column, plane} for the TMA descriptors
        double *fpool, *hpool;
        if ((rc = dev_alloc(c, S.allocations, &fpool, 3 * (size_t)L.plane))) return rc;
        if ((rc = dev_alloc(c, S.allocations, &hpool, 12 * (size_t)L.plane))) return rc;
        for (int k = 0; k < 3; k++) L.f[k] = fpool + k * L.plane;
        for (int k = 0; k < 5; k++) { L.FxH[k] = hpool + (2 * k) * L.plane; L.FpH[k] = hpool + (2 * k + 1) * L.plane; }
        L.FxL0 = hpool + 10 * L.plane; L.FpL0 = hpool + 11 * L.plane;
        if ((rc = dev_alloc(c, S.allocations, &L.chargeR, L.n_x))) return rc;
        if ((rc = dev_alloc(c, S.allocations, &L.currentR, L.n_x))) return rc;
        S.i_f0 = S.i_f1 = 0;
        if ((rc = vrt_fused_make_maps(c, s))) return rc;
        S.desc.assign(d, d + n_patches);       // only a fully built species carries descriptors
        return 0;
    }
    if ((rc = build_split(c, s, n_patches, d))) return rc;
    S.desc.assign(d, d + n_patches);
    return 0;
}

// Mesh::promoteHierarchyToMesh(false) without the host round trip: the new hierarchy's storage is
// built while the old one is still resident, Mesh::InterMeshDataTransfer (Mesh.cpp:116-130) runs as kernels between the two
// device patch tables, then the old storage is released.  The caller continues as after vrt_set_hierarchy + uploads:
// vrt_push_data, vrt_commit_state.
int vrt_regrid(vrt_ctx* c, int s, int n_patches, const vrt_patch_desc* d) {
    if (!c) return VRT_ERR_ARG;
    if (!check(c, c->grid_set && s >= 0 && s < c->n_species && c->S[s].configured && !c->S[s].desc.empty(), "vrt_regrid: no hierarchy to regrid")) return VRT_ERR_STATE;
    if (!check(c, n_patches >= 1 && d, "vrt_regrid: bad arguments")) return VRT_ERR_ARG;
    VrtSpeciesState& S = c->S[s];
    if (!check(c, S.path == VRT_PATH_SPLIT && c->max_depth >= 1, "vrt_regrid: only hierarchies on the split path regrid")) return VRT_ERR_STATE;
    const int r = c->refinement_ratio;
    // the coarse -> fine transfer fills a one-coarse-cell ring, i.e. r fine ghost cells, and the padded layout holds 2 (the
    // reference's own code carries the same restriction, Rectangle.cpp:892-918)
    if (!check(c, r == 2, "vrt_regrid: the regrid data movers support refinement ratio 2 only")) return VRT_ERR_ARG;
    for (int p = 0; p < n_patches; p++) {
        const vrt_patch_desc& q = d[p];
        if (!check(c, q.depth >= 0 && q.depth <= c->max_depth && q.n_x >= r && q.n_p >= r && q.n_x % r == 0 && q.n_p % r == 0,
                   "vrt_regrid: bad patch descriptor")) return VRT_ERR_ARG;
    }
    cudaSetDevice(c->device);
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    drop_graphs(c, true);
    // park the old storage; it comes back unchanged if the new hierarchy cannot be built or filled
    VrtSpeciesState old;
    auto exchange = [&]() {
        old.allocations.swap(S.allocations); old.patches.swap(S.patches); old.table.swap(S.table); old.table_index.swap(S.table_index);
        old.table_order.swap(S.table_order); old.level_patches.swap(S.level_patches); old.desc.swap(S.desc);
        std::swap(old.d_patches, S.d_patches); std::swap(old.conn_pool, S.conn_pool); std::swap(old.has_amr, S.has_amr);
        std::swap(old.d_lb, S.d_lb); old.lb_first.swap(S.lb_first); old.lb_count.swap(S.lb_count);
    };
    exchange();
    int rc;
    try { rc = build_split(c, s, n_patches, d); }
    catch (const std::exception&) { c->err = "vrt_regrid: out of host memory"; rc = VRT_ERR_NOMEM; }
    const long l0 = c->launches;
    if (!rc) rc = vrt_amr_transfer(c, old, S);
    if (getenv("VRT_TRACE")) fprintf(stderr, "vrt_regrid: species %d, %zu -> %d patches, %ld transfer kernels on the device\n", s, old.desc.size(), n_patches, c->launches - l0);
    if (!rc) { cudaError_t e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) { c->err = std::string("vrt_regrid: ") + cudaGetErrorString(e); rc = VRT_ERR_CUDA; } }
    if (rc) exchange();                       // the half-built new hierarchy is now in `old` and is released below
    else S.desc.assign(d, d + n_patches);
    for (double* p : old.allocations) cudaFree(p);
    if (old.d_patches) cudaFree(old.d_patches);
    if (old.conn_pool) cudaFree(old.conn_pool);
    return rc;
}

// Rectangle::ErrorEstimate (Rectangle.hpp:128-130) > refinementCriteria for every interior cell of a patch, as a byte map
// (n_x x n_p, p fast); the caller applies the margins of Rectangle::getError (Rectangle.cpp:866-890) and ORs the user's
// RefinementOverride.  1 byte per cell crosses the bus instead of the 16 bytes of the two f states.
int vrt_error_flags(vrt_ctx* c, int s, int patch, const double weights[5], double criteria, unsigned char* flags_host) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, weights && flags_host && patch >= 0 && patch < (int)c->S[s].desc.size(), "vrt_error_flags: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->S[s].path == VRT_PATH_SPLIT, "vrt_error_flags: split path only")) return VRT_ERR_STATE;
    return vrt_amr_error_flags(c, s, patch, weights, criteria, flags_host);
}

// ---- data movement ----------------------------------------------------------------------------------
static int slab_plane_index(const VrtSpeciesState& S, int state) {
    if (state == 0) return S.i_f0;
    if (state == 1) return S.i_f1;
    return -1;
}

int vrt_patch_upload_f(vrt_ctx* c, int s, int patch, int state, const double* host) {
    if (!c || !host) return VRT_ERR_ARG;
    if (!check(c, s >= 0 && s < c->n_species && patch >= 0 && patch < (int)c->S[s].desc.size() && state >= 0 && state <= 2, "vrt_patch_upload_f: bad arguments")) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    VrtSpeciesState& S = c->S[s];
    if (S.path == VRT_PATH_FUSED) {
        if (!check(c, state != 2, "vrt_patch_upload_f: the fused path keeps no predictor state")) return VRT_ERR_ARG;
        VrtSlabDev& L = S.slab;
        // uploading state 1 while both states alias one plane splits them
        if (state == 1 && S.i_f1 == S.i_f0) {
            S.i_f1 = (S.i_f0 + 1) % 3;
        }
        int pi = slab_plane_index(S, state);
        // host: (n_xg+4) x (n_p+4), cell (i,j) at (n_p+4)*(i+2)+2+j.  device column c (local) <- host column x_begin+c.
        // halo columns inside the global domain come from the host array; beyond its 2 ghost columns they stay 0.
        const int hp = L.n_p + 4;
        for (int cl = -L.gx; cl < L.n_x + L.gx; cl++) {
            int gi = L.x_begin + cl;
            if (gi < -2 || gi >= L.n_x_global + 2) continue;
            VRT_CUDA(c, cudaMemcpyAsync(L.f[pi] + (long)(cl + L.gx) * L.pitch + (VRT_SLAB_GH - 2), host + (long)(gi + 2) * hp, sizeof(double) * hp,
                                        cudaMemcpyHostToDevice, c->stream));
        }
        VRT_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    VrtPatchDev& P = S.patches[patch];
    double* dst = state == 0 ? P.f0 : (state == 1 ? P.f1 : P.f2);
    VRT_CUDA(c, cudaMemcpyAsync(dst, host, sizeof(double) * P.npad, cudaMemcpyHostToDevice, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int vrt_patch_download_f(vrt_ctx* c, int s, int patch, int state, double* host) {
    if (!c || !host) return VRT_ERR_ARG;
    if (!check(c, s >= 0 && s < c->n_species && patch >= 0 && patch < (int)c->S[s].desc.size() && state >= 0 && state <= 2, "vrt_patch_download_f: bad arguments")) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    VrtSpeciesState& S = c->S[s];
    if (S.path == VRT_PATH_FUSED) {
        if (!check(c, state != 2, "vrt_patch_download_f: the fused path keeps no predictor state")) return VRT_ERR_ARG;
        VrtSlabDev& L = S.slab;
        int pi = slab_plane_index(S, state);
        const int hp = L.n_p + 4;
        // only this slab's own columns (plus the physical ghost columns it touches) are written to the host array
        int lo = L.left ? -2 : 0, hi = L.right ? L.n_x + 2 : L.n_x;
        VRT_CUDA(c, cudaMemcpy2DAsync(host + (long)(L.x_begin + lo + 2) * hp, sizeof(double) * hp,
                                      L.f[pi] + (long)(lo + L.gx) * L.pitch + (VRT_SLAB_GH - 2), sizeof(double) * L.pitch, sizeof(double) * hp, hi - lo,
                                      cudaMemcpyDeviceToHost, c->stream));
        VRT_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    VrtPatchDev& P = S.patches[patch];
    const double* src = state == 0 ? P.f0 : (state == 1 ? P.f1 : P.f2);
    VRT_CUDA(c, cudaMemcpyAsync(host, src, sizeof(double) * P.npad, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int vrt_commit_state(vrt_ctx* c, int s) {
    if (!c || s < 0 || s >= c->n_species) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    VrtSpeciesState& S = c->S[s];
    if (S.path == VRT_PATH_FUSED) { S.i_f0 = S.i_f1; return 0; }
    for (size_t d = 0; d < S.level_patches.size(); d++)
        if (int r = vrt_split_substep(c, s, (int)d, &c->d_params->dt, 5, 3)) return r;
    return 0;
}

int vrt_patch_download_plane(vrt_ctx* c, int s, int patch, int which, int slot, double* host) {
    if (!c || !host) return VRT_ERR_ARG;
    if (!check(c, s >= 0 && s < c->n_species && patch >= 0 && patch < (int)c->S[s].desc.size() && slot >= 0 && slot < 6, "vrt_patch_download_plane: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->S[s].path == VRT_PATH_SPLIT, "vrt_patch_download_plane: work planes exist on the split path only")) return VRT_ERR_STATE;
    cudaSetDevice(c->device);
    const VrtPatchDev& P = c->S[s].patches[patch];
    const double* src = nullptr;
    switch (which) {
        case VRT_PLANE_FXH: src = P.FxH + (long)slot * P.npad; break;
        case VRT_PLANE_FPH: src = P.FpH + (long)slot * P.npad; break;
        case VRT_PLANE_FXL: src = P.FxL; break;
        case VRT_PLANE_FPL: src = P.FpL; break;
        case VRT_PLANE_FXDS: src = P.FxDS; break;
        case VRT_PLANE_FPDS: src = P.FpDS; break;
        case VRT_PLANE_RP: src = P.Rp; break;
        case VRT_PLANE_RM: src = P.Rm; break;
        case VRT_PLANE_CX: src = P.Cx; break;
        case VRT_PLANE_CP: src = P.Cp; break;
        case VRT_PLANE_EX: src = P.ex; break;
        case VRT_PLANE_EP: src = P.ep; break;
        case VRT_PLANE_FX: src = P.fx; break;
        case VRT_PLANE_FP: src = P.fp; break;
        default: c->err = "vrt_patch_download_plane: bad selector"; return VRT_ERR_ARG;
    }
    VRT_CUDA(c, cudaMemcpyAsync(host, src, sizeof(double) * P.npad, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int vrt_field_upload(vrt_ctx* c, int which, int slot, const double* host) {
    if (!c || !host || which < 0 || which > 5 || slot < 0 || slot > 7 || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    VRT_CUDA(c, cudaMemcpyAsync(c->F.Y[which] + (long)slot * c->F.M, host, sizeof(double) * c->F.M, cudaMemcpyHostToDevice, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vrt_field_download(vrt_ctx* c, int which, int slot, double* host) {
    if (!c || !host || which < 0 || which > 5 || slot < 0 || slot > 7 || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    VRT_CUDA(c, cudaMemcpyAsync(host, c->F.Y[which] + (long)slot * c->F.M, sizeof(double) * c->F.M, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

static double* sel_1d(vrt_ctx* c, int which, long* n) {
    VrtFields& F = c->F;
    *n = F.N;
    switch (which) {
        case VRT_PHI: return F.PHI;
        case VRT_CHARGE: return F.charge;
        case VRT_J: return F.J;
        case VRT_A_SQUARED: *n = F.N + 1; return F.a_squared;
        case VRT_NEUTRALIZATION: return F.neutral;
        case VRT_EFIELD: return F.E + F.epad;
        default:
            if (which >= VRT_CHARGES0 && which < VRT_CHARGES0 + c->n_species) return c->S[which - VRT_CHARGES0].d_charges;
    }
    return nullptr;
}
int vrt_set_1d(vrt_ctx* c, int which, const double* host) {
    if (!c || !host || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    long n; double* d = sel_1d(c, which, &n);
    if (!check(c, d != nullptr && which != VRT_EFIELD, "vrt_set_1d: bad selector")) return VRT_ERR_ARG;
    VRT_CUDA(c, cudaMemcpyAsync(d, host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (which == VRT_PHI) return vrt_fields_refresh_efield(c);
    return 0;
}
int vrt_get_1d(vrt_ctx* c, int which, double* host) {
    if (!c || !host || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    long n; double* d = sel_1d(c, which, &n);
    if (!check(c, d != nullptr, "vrt_get_1d: bad selector")) return VRT_ERR_ARG;
    VRT_CUDA(c, cudaMemcpyAsync(host, d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int vrt_set_scalar(vrt_ctx* c, int which, double v) {
    if (!c || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    if (which == VRT_TIME) { c->time = v; return 0; }
    if (which == VRT_EX0) {
        k_fill<<<1, 1, 0, c->stream>>>(c->F.Ex0, 1, v);
        VRT_CUDA(c, cudaGetLastError());
        return vrt_fields_refresh_efield(c);
    }
    return VRT_ERR_ARG;
}
int vrt_get_scalar(vrt_ctx* c, int which, double* v) {
    if (!c || !v || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    if (which == VRT_TIME) { *v = c->time; return 0; }
    if (which == VRT_EX0) {
        VRT_CUDA(c, cudaMemcpyAsync(v, c->F.Ex0, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        VRT_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    return VRT_ERR_ARG;
}

// ---- hot path ---------------------------------------------------------------------------------------
static int ready(vrt_ctx* c) {
    if (!c) return VRT_ERR_ARG;
    if (!c->grid_set) { c->err = "grid not set"; return VRT_ERR_STATE; }
    for (auto& S : c->S) if (S.desc.empty()) { c->err = "hierarchy not set for every species"; return VRT_ERR_STATE; }
    cudaSetDevice(c->device);
    return 0;
}

// per-species calls only need that species' hierarchy (SolverManager constructs the meshes one after the other)
static int ready_species(vrt_ctx* c, int s) {
    if (!c) return VRT_ERR_ARG;
    if (!c->grid_set) { c->err = "grid not set"; return VRT_ERR_STATE; }
    if (s < 0 || s >= c->n_species) { c->err = "bad species index"; return VRT_ERR_ARG; }
    if (c->S[s].desc.empty()) { c->err = "hierarchy not set for this species"; return VRT_ERR_STATE; }
    cudaSetDevice(c->device);
    return 0;
}

static int species_moment_kernel(vrt_ctx* c, int s) {
    return (c->S[s].path == VRT_PATH_FUSED) ? vrt_fused_moments(c, s) : vrt_split_moments(c, s);
}

// fn(s) for every species: species s > 0 on its own stream between a fork and a join event (concurrent branches of the step graph),
// species 0 on the context's stream.  fn enqueues on c->stream.
extern "C++" {
template <typename Fn>
static int for_each_species_forked(vrt_ctx* c, Fn fn) {
    int r;
    const bool fork = c->fork_species && c->n_species > 1;
    if (!fork) {
        for (int s = 0; s < c->n_species; s++) if ((r = fn(s))) return r;
        return 0;
    }
    if (!c->ev_fork) {
        VRT_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        c->aux_stream.resize(c->n_species - 1); c->aux_join.resize(c->n_species - 1);
        for (int s = 1; s < c->n_species; s++) {
            VRT_CUDA(c, cudaStreamCreateWithFlags(&c->aux_stream[s - 1], cudaStreamNonBlocking));
            VRT_CUDA(c, cudaEventCreateWithFlags(&c->aux_join[s - 1], cudaEventDisableTiming));
        }
    }
    cudaStream_t main_stream = c->stream;
    VRT_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
    for (int s = 1; s < c->n_species; s++) {
        cudaStream_t aux = c->aux_stream[s - 1];
        VRT_CUDA(c, cudaStreamWaitEvent(aux, c->ev_fork, 0));
        c->stream = aux;
        r = fn(s);
        c->stream = main_stream;
        if (r) return r;
        VRT_CUDA(c, cudaEventRecord(c->aux_join[s - 1], aux));
    }
    if ((r = fn(0))) return r;
    for (int s = 1; s < c->n_species; s++) VRT_CUDA(c, cudaStreamWaitEvent(main_stream, c->aux_join[s - 1], 0));
    return 0;
}
}  // extern "C++"

// EMFieldSolver::AssembleRhoAndJ (EMSolver.cpp:104-122): the species' moment kernels on concurrent branches (they only write their own
// per-patch / per-slab arrays), then ONE assembly launch that forms charges[s], J and the total charge in the reference's order of
// additions; on x-slab runs the slabs are gathered in between and the total follows the gather
static int moments_impl(vrt_ctx* c) {
    int r;
    if ((r = for_each_species_forked(c, [&](int s) { return species_moment_kernel(c, s); }))) return r;
    const unsigned all = (1u << c->n_species) - 1u;
    if ((r = vrt_fields_assemble(c, all, c->n_ranks == 1))) return r;
    if (c->n_ranks > 1) {
        if ((r = vrt_comm_gather_moments(c))) return r;
        if ((r = vrt_fields_total_charge(c))) return r;
    }
    return 0;
}

int vrt_moments(vrt_ctx* c) { if (int r = ready(c)) return r; return moments_impl(c); }

int vrt_moments_species(vrt_ctx* c, int s, double* charge_host, double* j_host) {
    if (int r = ready(c)) return r;
    if (!check(c, s >= 0 && s < c->n_species && charge_host && j_host, "vrt_moments_species: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->n_ranks == 1, "vrt_moments_species: single-rank contexts only")) return VRT_ERR_STATE;
    // the species' contribution alone: charges[s] and the total current are rebuilt from this species only, copied out,
    // and the assembled state (all species) is restored by a full vrt_moments afterwards
    int r;
    if ((r = species_moment_kernel(c, s))) return r;
    if ((r = vrt_fields_assemble(c, 1u << s, 0))) return r;
    const int N = c->F.N;
    std::vector<double> a(N), b(N);
    VRT_CUDA(c, cudaMemcpyAsync(a.data(), c->S[s].d_charges, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaMemcpyAsync(b.data(), c->F.J, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < N; i++) { charge_host[i] += a[i]; j_host[i] += b[i]; }
    return moments_impl(c);
}

// Rectangle::chargeR / currentR of one patch after Rectangle::CalculateRhoAndJ (what Level::CollectRhoAndJ reads, Level.cpp:42-62)
int vrt_patch_moments(vrt_ctx* c, int s, int patch, double* charge_r_host, double* current_r_host) {
    if (int r = ready(c)) return r;          // the closing vrt_moments needs every species' hierarchy
    if (!check(c, s >= 0 && s < c->n_species && charge_r_host && current_r_host && patch >= 0 && patch < (int)c->S[s].desc.size(),
               "vrt_patch_moments: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->n_ranks == 1, "vrt_patch_moments: single-rank contexts only")) return VRT_ERR_STATE;
    VrtSpeciesState& S = c->S[s];
    // the species' moment kernels also rebuild charges[s] and J from this species alone (as in vrt_moments_species); the
    // assembled state of all species is restored by a full vrt_moments afterwards
    int r;
    if ((r = species_moment_kernel(c, s))) return r;
    const double *d_charge, *d_current;
    size_t n;
    if (S.path == VRT_PATH_FUSED) { d_charge = S.slab.chargeR; d_current = S.slab.currentR; n = (size_t)S.slab.n_x; }
    else { const VrtPatchDev& P = S.patches[patch]; d_charge = P.chargeR; d_current = P.currentR; n = (size_t)P.n_x * P.rtb; }
    VRT_CUDA(c, cudaMemcpyAsync(charge_r_host, d_charge, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaMemcpyAsync(current_r_host, d_current, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return moments_impl(c);
}

// Rectangle::CalculateEnergy (Rectangle.cpp:284-305): the patch's contribution to dN/dp on the finest p grid
int vrt_patch_energy(vrt_ctx* c, int s, int patch, double* energy_host) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, energy_host && patch >= 0 && patch < (int)c->S[s].desc.size(), "vrt_patch_energy: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->n_ranks == 1, "vrt_patch_energy: single-rank contexts only")) return VRT_ERR_STATE;
    return vrt_split_patch_energy(c, s, patch, energy_host);
}

int vrt_enforce_neutralization(vrt_ctx* c) {
    if (int r = ready(c)) return r;
    if (int r = moments_impl(c)) return r;
    return vrt_fields_neutralize(c);
}

int vrt_poisson(vrt_ctx* c) { if (int r = ready(c)) return r; return vrt_fields_poisson(c); }

static int push_data_impl(vrt_ctx* c, int s, int val) {
    VrtSpeciesState& S = c->S[s];
    if (S.path == VRT_PATH_FUSED) return 0;   // ghosts of the slab planes are never written; halos are exchanged per stage
    return vrt_amr_push_data(c, s, val);   // Mesh::PushData (Mesh.cpp:91-106)
}

static int vlasov_stage_impl(vrt_ctx* c, int s, const double* d_dt, int step) {
    VrtSpeciesState& S = c->S[s];
    int r;
    if (S.path == VRT_PATH_FUSED) {
        // x-slabs: the kernel reads the halo columns the previous exchange of this species delivered; its own exchange runs
        // on the communication stream and overlaps whatever the compute stream does next
        if (c->n_ranks > 1 && (r = vrt_comm_wait_halo(c, s))) return r;
        if ((r = vrt_fused_stage(c, s, d_dt, step))) return r;
        if (c->n_ranks > 1 && (r = vrt_comm_halo_exchange(c, s))) return r;
        return 0;
    }
    // Mesh::Advance (Mesh.cpp:64-89); each sub-step serves all levels in one launch (vrt_split_substep_all)
    if ((r = vrt_split_substep_all(c, s, d_dt, step, 0))) return r;
    if ((r = push_data_impl(c, s, 2))) return r;
    if ((r = vrt_split_substep_all(c, s, d_dt, step, 1))) return r;
    if ((r = vrt_amr_push_boundary_c(c, s))) return r;
    if ((r = vrt_split_substep_all(c, s, d_dt, step, 2))) return r;
    if ((r = push_data_impl(c, s, 1))) return r;
    if (step == 5 && (r = vrt_split_substep_all(c, s, d_dt, step, 3))) return r;
    return 0;
}

int vrt_vlasov_stage(vrt_ctx* c, int s, double dt, int step) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species && step >= 0 && step <= 5, "vrt_vlasov_stage: bad arguments")) return VRT_ERR_ARG;
    if (int r = set_params_async(c, dt, nullptr)) return r;
    if (int r = vlasov_stage_impl(c, s, &c->d_params->dt, step)) return r;
    return c->n_ranks > 1 ? vrt_comm_wait_halo(c, s) : 0;
}

int vrt_vlasov_substep(vrt_ctx* c, int s, int depth, double dt, int step, int substep) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species && step >= -1 && step <= 5, "vrt_vlasov_substep: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->S[s].path == VRT_PATH_SPLIT, "vrt_vlasov_substep: sub-steps exist on the split path only (vrt_set_path)")) return VRT_ERR_STATE;
    if (int r = set_params_async(c, dt, nullptr)) return r;
    return vrt_split_substep(c, s, depth, &c->d_params->dt, step < 0 ? 0 : step, substep);
}

int vrt_push_data(vrt_ctx* c, int s, int val) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species && (val == 1 || val == 2), "vrt_push_data: bad arguments")) return VRT_ERR_ARG;
    return push_data_impl(c, s, val);
}
int vrt_level_push(vrt_ctx* c, int s, int depth, int update_type, int val) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species && (val == 1 || val == 2 || update_type >= 4), "vrt_level_push: bad arguments")) return VRT_ERR_ARG;
    if (!check(c, c->S[s].path == VRT_PATH_SPLIT, "vrt_level_push: per-level passes exist on the split path only")) return VRT_ERR_STATE;
    return vrt_amr_level_pass(c, s, depth, update_type, val);
}
int vrt_push_boundary_c(vrt_ctx* c, int s) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species, "vrt_push_boundary_c: bad arguments")) return VRT_ERR_ARG;
    if (c->S[s].path == VRT_PATH_FUSED) return 0;   // the fused kernel carries the limiter across its own tiles
    return vrt_amr_push_boundary_c(c, s);
}

int vrt_field_stage(vrt_ctx* c, int step, double dt, double by0, double bz0) {
    if (!c || !c->grid_set) return VRT_ERR_ARG;
    if (!check(c, step >= 0 && step <= 5, "vrt_field_stage: bad step")) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    double laser[12] = {0};
    laser[2 * step] = by0; laser[2 * step + 1] = bz0;
    if (int r = set_params_async(c, dt, laser)) return r;
    return vrt_fields_rhs_update_faces(c, step, c->d_params);
}

int vrt_cfl_bound(vrt_ctx* c, double* out) {
    if (!c || !out || !c->grid_set) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    if (int r = vrt_fields_cfl(c)) return r;
    VRT_CUDA(c, cudaMemcpyAsync(out, c->F.cfl + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

double vrt_update_time(double time, int step, double dt) {   // Settings::UpdateTime (Settings.cpp:166-179)
    if (step == 1) time += (0.5 * dt);
    else if (step == 2) time += (0.332 - 0.5) * dt;
    else if (step == 3) time += (0.62 - 0.332) * dt;
    else if (step == 4) time += (0.85 - 0.62) * dt;
    else if (step == 5) time += (1.0 - 0.85) * dt;
    return time;
}

// Mesh::Advance of every species for one RK stage (SolverManager.cpp:33-35).  The species do not interact inside a stage, so
// species s > 0 is enqueued on its own stream between a fork and a join event: concurrent branches of the step graph.  This
// halves the launch-latency-bound time of small and AMR hierarchies and fills the tail of the last wave of large kernels.
static int vlasov_stages_all(vrt_ctx* c, int i) {
    // (x-slab runs too: the NCCL calls all go to the communication stream, in host order, whichever stream computes)
    return for_each_species_forked(c, [&](int s) { return vlasov_stage_impl(c, s, &c->d_params->dt, i); });
}

// the six stages of SolverManager::Advance as stream work (SolverManager.cpp:28-39)
static int enqueue_step(vrt_ctx* c) {
    int r;
    const bool fork = c->fork_fields;
    if (fork && !c->field_stream) {
        VRT_CUDA(c, cudaStreamCreateWithFlags(&c->field_stream, cudaStreamNonBlocking));
        VRT_CUDA(c, cudaEventCreateWithFlags(&c->ev_ffork, cudaEventDisableTiming));
        VRT_CUDA(c, cudaEventCreateWithFlags(&c->ev_fjoin, cudaEventDisableTiming));
    }
    for (int i = 0; i < 6; i++) {
        if ((r = moments_impl(c))) return r;
        if (fork) {
            // RGKStep(i) reads J and the fields only (EMSolver.cpp:479-553): a side branch next to UpdatePotential and Mesh::Advance,
            // which read the a^2 and E of the stage's start; joined before the next stage's moments read the new A
            cudaStream_t main_stream = c->stream;
            VRT_CUDA(c, cudaEventRecord(c->ev_ffork, main_stream));
            VRT_CUDA(c, cudaStreamWaitEvent(c->field_stream, c->ev_ffork, 0));
            c->stream = c->field_stream;
            r = vrt_fields_rhs_update_faces(c, i, c->d_params, c->asq_alt);
            c->stream = main_stream;
            if (r) return r;
            VRT_CUDA(c, cudaEventRecord(c->ev_fjoin, c->field_stream));
        }
        if ((r = vrt_fields_poisson(c))) return r;
        if ((r = vlasov_stages_all(c, i))) return r;
        if (fork) {
            VRT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_fjoin, 0));
            std::swap(c->F.a_squared, c->asq_alt);         // six exchanges per step: every step starts on the same buffer
        } else if ((r = vrt_fields_rhs_update_faces(c, i, c->d_params))) return r;
    }
    if (c->n_ranks > 1 && (r = vrt_comm_wait_halo(c, -1))) return r;      // a finished step has its halos in place
    return 0;
}

int vrt_step(vrt_ctx* c, double dt, const double laser[12]) {
    if (int r = ready(c)) return r;
    if (!check(c, laser != nullptr, "vrt_step: laser values missing")) return VRT_ERR_ARG;
    if (int r = set_params_async(c, dt, laser)) return r;
    const long l0 = c->launches;
    // The plane rotation of the fused path has period 2 after the first step; graphs are cached per rotation state.
    int key = 0;
    for (auto& S : c->S) if (S.path == VRT_PATH_FUSED) { key = S.i_f0; if (!check(c, S.i_f0 == S.i_f1, "vrt_step: mid-step state")) return VRT_ERR_STATE; }
    for (auto& S : c->S) if (S.path == VRT_PATH_FUSED && !check(c, S.i_f0 == key, "vrt_step: species out of phase")) return VRT_ERR_STATE;
    // x-slab runs: the step's NCCL calls (grouped send/recv, all-gather) sit on the communication stream between events of the
    // compute streams and are captured with it (NCCL >= 2.9 is capturable; every rank captures the same sequence).  VRT_MULTI_GRAPH=0
    // issues the step eagerly instead.
    static const bool multi_graph = !(getenv("VRT_MULTI_GRAPH") && atoi(getenv("VRT_MULTI_GRAPH")) == 0);
    const bool use_graph = c->use_graph && (c->n_ranks == 1 || multi_graph);
    if (!use_graph) {
        if (int r = enqueue_step(c)) return r;
        c->last_step_launches = c->launches - l0;
    } else {
        if (!c->graph_step3[key]) {
            std::vector<std::pair<int, int>> saved;
            for (auto& S : c->S) saved.push_back({S.i_f0, S.i_f1});
            VRT_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int r = enqueue_step(c);
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            c->graph_launches[key] = c->launches - l0;
            for (size_t s = 0; s < c->S.size(); s++) { c->graph_end_state[key][s] = {c->S[s].i_f0, c->S[s].i_f1}; c->S[s].i_f0 = saved[s].first; c->S[s].i_f1 = saved[s].second; }
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) { c->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
            if (c->graph_stale[key]) {       // after a regrid: same topology -> update the parked executable in place (VRT_GRAPH_UPDATE=0: A/B switch)
                static const bool update_on = !(getenv("VRT_GRAPH_UPDATE") && atoi(getenv("VRT_GRAPH_UPDATE")) == 0);
                cudaGraphExecUpdateResultInfo info;
                if (update_on && cudaGraphExecUpdate(c->graph_stale[key], g, &info) == cudaSuccess) c->graph_step3[key] = c->graph_stale[key];
                else { cudaGetLastError(); cudaGraphExecDestroy(c->graph_stale[key]); }
                c->graph_stale[key] = nullptr;
            }
            if (!c->graph_step3[key]) {
                e = cudaGraphInstantiate(&c->graph_step3[key], g, 0);
                if (e != cudaSuccess) { cudaGraphDestroy(g); c->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
            }
            cudaGraphDestroy(g);
        }
        VRT_CUDA(c, cudaGraphLaunch(c->graph_step3[key], c->stream));
        for (size_t s = 0; s < c->S.size(); s++) { c->S[s].i_f0 = c->graph_end_state[key][s].first; c->S[s].i_f1 = c->graph_end_state[key][s].second; }
        c->last_step_launches = c->graph_launches[key];
    }
    for (int i = 0; i < 6; i++) c->time = vrt_update_time(c->time, i, dt);
    return 0;
}

int vrt_step_fields(vrt_ctx* c, double dt, const double laser[12]) {
    if (!c || !c->grid_set || !laser) return VRT_ERR_ARG;
    cudaSetDevice(c->device);
    if (int r = set_params_async(c, dt, laser)) return r;
    const long l0 = c->launches;
    // the 18 launches of the six field stages as one graph (the fields-only phase of a fine mesh is tens of thousands of steps,
    // veritas.cpp:135-144, each a chain of latency-bound 1-D kernels); dt and the laser values are read from the parameter block
    if (c->use_graph) {
        if (!c->graph_fields) {
            VRT_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            int r = 0;
            for (int i = 0; i < 6 && !r; i++) r = vrt_fields_rhs_update_faces(c, i, c->d_params);
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            c->graph_fields_launches = c->launches - l0;
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) { c->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
            e = cudaGraphInstantiate(&c->graph_fields, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { c->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return VRT_ERR_CUDA; }
        }
        VRT_CUDA(c, cudaGraphLaunch(c->graph_fields, c->stream));
        c->last_step_launches = c->graph_fields_launches;
    } else {
        for (int i = 0; i < 6; i++) if (int r = vrt_fields_rhs_update_faces(c, i, c->d_params)) return r;
        c->last_step_launches = c->launches - l0;
    }
    for (int i = 0; i < 6; i++) c->time = vrt_update_time(c->time, i, dt);
    return 0;
}

long vrt_last_step_launches(const vrt_ctx* c) { return c ? c->last_step_launches : 0; }

int vrt_get_hierarchy(vrt_ctx* c, int s, int capacity, vrt_patch_desc* out) {
    if (int r = ready_species(c, s)) return r;
    const std::vector<vrt_patch_desc>& d = c->S[s].desc;
    if (out) for (int k = 0; k < capacity && k < (int)d.size(); k++) out[k] = d[k];
    return (int)d.size();
}

int vrt_fused_plan(vrt_ctx* c, int s, int out[6]) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, out && c->S[s].path == VRT_PATH_FUSED, "vrt_fused_plan: species is not on the fused path")) return VRT_ERR_STATE;
    return vrt_fused_plan_impl(c, s, out);
}

int vrt_set_option(vrt_ctx* c, int option, int value) {
    if (!c) return VRT_ERR_ARG;
    if (option == 0) { c->use_graph = value != 0; return 0; }
    if (option == 1) { c->fork_species = value != 0; drop_graphs(c); return 0; }
    if (option == 2) { c->fork_fields = value != 0; drop_graphs(c); return 0; }
    return VRT_ERR_ARG;
}

int vrt_init_maxwellian_slab(vrt_ctx* c, int s, double xl, double xr, double n0, double T, int quadrature_depth) {
    if (int r = ready_species(c, s)) return r;
    if (!check(c, s >= 0 && s < c->n_species, "vrt_init_maxwellian_slab: bad species")) return VRT_ERR_ARG;
    return vrt_init_kernels_maxwellian(c, s, xl, xr, n0, T, quadrature_depth);
}

}  // extern "C"
