// Binary checkpoint / restart of a context at a step boundary (the reference has none: its only
// persistent output are lossy text dumps, Mesh.cpp:889, EMSolver.cpp:344).  A restart continues bit for bit.
//
// What a step boundary needs (everything else is rebuilt before it is read): per species the hierarchy and f
// (f^n = stage value there: Rectangle.cpp:1614-1622); the six transverse field arrays with their 8 slots (EMSolver.hpp:10-23);
// a_squared; PHI and the incrementally updated Ex0 (quirk Q4), from which the E table CalculateDt reads is refreshed; the
// neutralisation charge; Settings::time.  Each rank of an x-slab run writes its own file (its slab planes with the exchanged
// halo columns).
#include "vrt_internal.cuh"
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

int vrt_fields_refresh_efield(vrt_ctx* c);

namespace {

constexpr char kMagic[8] = {'V', 'R', 'T', 'C', 'K', 'P', 'T', '1'};
constexpr size_t kChunk = 1 << 22;      // doubles per staging transfer (32 MiB, pinned)

struct Header {
    char magic[8];
    int n_species, N, pre, post, r, max_depth, rank, n_ranks, x_begin, x_end;
    double dx, time;
};

struct Stager {
    vrt_ctx* c; FILE* f; double* pinned = nullptr; std::string err;
    Stager(vrt_ctx* c_, FILE* f_) : c(c_), f(f_) { if (cudaMallocHost(&pinned, kChunk * sizeof(double)) != cudaSuccess) pinned = nullptr; }
    ~Stager() { if (pinned) cudaFreeHost(pinned); }
    bool ok() const { return pinned != nullptr; }
    bool write(const double* dev, size_t n) {
        for (size_t o = 0; o < n; o += kChunk) {
            const size_t m = std::min(kChunk, n - o);
            if (cudaMemcpyAsync(pinned, dev + o, m * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return false;
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
            if (fwrite(pinned, sizeof(double), m, f) != m) return false;
        }
        return true;
    }
    bool read(double* dev, size_t n) {
        for (size_t o = 0; o < n; o += kChunk) {
            const size_t m = std::min(kChunk, n - o);
            if (fread(pinned, sizeof(double), m, f) != m) return false;
            if (cudaMemcpyAsync(dev + o, pinned, m * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return false;
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
        }
        return true;
    }
};

int fail(vrt_ctx* c, FILE* f, const std::string& msg, int code) { if (f) fclose(f); c->err = msg; return code; }

}  // namespace

extern "C" {

int vrt_checkpoint_write(vrt_ctx* c, const char* path) {
    if (!c || !path) return VRT_ERR_ARG;
    if (!c->grid_set) { c->err = "vrt_checkpoint_write: grid not set"; return VRT_ERR_STATE; }
    for (auto& S : c->S) {
        if (S.desc.empty()) { c->err = "vrt_checkpoint_write: hierarchy not set for every species"; return VRT_ERR_STATE; }
        if (S.path == VRT_PATH_FUSED && S.i_f0 != S.i_f1) { c->err = "vrt_checkpoint_write: not at a step boundary"; return VRT_ERR_STATE; }
    }
    cudaSetDevice(c->device);
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    FILE* f = fopen(path, "wb");
    if (!f) return fail(c, nullptr, std::string("vrt_checkpoint_write: cannot open ") + path, VRT_ERR_ARG);
    const VrtFields& F = c->F;
    Header h{};
    std::memcpy(h.magic, kMagic, 8);
    h.n_species = c->n_species; h.N = F.N; h.pre = F.pre; h.post = F.post; h.r = c->refinement_ratio; h.max_depth = c->max_depth;
    h.rank = c->rank; h.n_ranks = c->n_ranks; h.x_begin = c->x_begin; h.x_end = c->x_end; h.dx = F.dx; h.time = c->time;
    if (fwrite(&h, sizeof(h), 1, f) != 1) return fail(c, f, "vrt_checkpoint_write: write failed", VRT_ERR_CUDA);
    Stager st(c, f);
    if (!st.ok()) return fail(c, f, "vrt_checkpoint_write: no pinned staging buffer", VRT_ERR_NOMEM);
    bool ok = true;
    for (int v = 0; v < 6 && ok; v++) ok = st.write(F.Y[v], 8L * F.M);
    ok = ok && st.write(F.a_squared, F.N + 1) && st.write(F.PHI, F.N) && st.write(F.neutral, F.N) && st.write(F.Ex0, 1);
    for (int s = 0; s < c->n_species && ok; s++) {
        const VrtSpeciesState& S = c->S[s];
        const int n = (int)S.desc.size();
        ok = fwrite(&S.sp, sizeof(VrtSpecies), 1, f) == 1 && fwrite(&S.path, sizeof(int), 1, f) == 1 && fwrite(&n, sizeof(int), 1, f) == 1 &&
             fwrite(S.desc.data(), sizeof(vrt_patch_desc), n, f) == (size_t)n;
        if (!ok) break;
        if (S.path == VRT_PATH_FUSED) ok = st.write(S.slab.f[S.i_f0], (size_t)S.slab.plane);
        else for (int p = 0; p < n && ok; p++) ok = st.write(S.patches[p].f1, (size_t)S.patches[p].npad);
    }
    if (!ok) return fail(c, f, "vrt_checkpoint_write: write failed (disk full?)", VRT_ERR_CUDA);
    fclose(f);
    return 0;
}

// Restore.  Transactional: pass 1 walks the whole file — header, every species record (species constants, path, a bounded patch
// count, the descriptors, which must form a hierarchy vrt_set_hierarchy accepts on the path the writer used) and the size of
// every data block, which must add up to the file's length — without touching the context.  Only then pass 2 overwrites the
// field arrays and rebuilds the hierarchies.  If pass 2 still fails (device out of memory, I/O error), the context is left without
// hierarchies, so that every later hot-path call answers VRT_ERR_STATE instead of running on a half-restored state.
static int checkpoint_read_impl(vrt_ctx* c, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(c, nullptr, std::string("vrt_checkpoint_read: cannot open ") + path, VRT_ERR_ARG);
    Header h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, kMagic, 8)) return fail(c, f, "vrt_checkpoint_read: not a veritas_b200 checkpoint", VRT_ERR_ARG);
    VrtFields& F = c->F;
    if (h.n_species != c->n_species || h.N != F.N || h.pre != F.pre || h.post != F.post || h.r != c->refinement_ratio || h.max_depth != c->max_depth ||
        h.rank != c->rank || h.n_ranks != c->n_ranks || h.x_begin != c->x_begin || h.x_end != c->x_end || h.dx != F.dx)
        return fail(c, f, "vrt_checkpoint_read: the checkpoint was written for another grid / decomposition", VRT_ERR_STATE);
    if (fseek(f, 0, SEEK_END)) return fail(c, f, "vrt_checkpoint_read: cannot seek", VRT_ERR_ARG);
    const long file_size = ftell(f);
    const long fields_doubles = 6 * 8L * F.M + (F.N + 1) + F.N + F.N + 1;      // Y[6][8][M], a_squared, PHI, neutral, Ex0
    // ---- pass 1: validate everything, remember where each species record starts --------------------------------------------
    struct Rec { VrtSpecies sp; int path; std::vector<vrt_patch_desc> d; long data_off; };
    std::vector<Rec> recs(c->n_species);
    long off = (long)sizeof(Header) + fields_doubles * (long)sizeof(double);
    for (int s = 0; s < c->n_species; s++) {
        Rec& R = recs[s];
        int n = 0;
        if (off + (long)(sizeof(VrtSpecies) + 2 * sizeof(int)) > file_size || fseek(f, off, SEEK_SET) ||
            fread(&R.sp, sizeof(VrtSpecies), 1, f) != 1 || fread(&R.path, sizeof(int), 1, f) != 1 || fread(&n, sizeof(int), 1, f) != 1)
            return fail(c, f, "vrt_checkpoint_read: truncated file", VRT_ERR_ARG);
        off += (long)(sizeof(VrtSpecies) + 2 * sizeof(int));
        if (n < 1 || n > 65535 || (long)n * (long)sizeof(vrt_patch_desc) > file_size - off)
            return fail(c, f, "vrt_checkpoint_read: corrupt patch count", VRT_ERR_ARG);
        if (!(R.sp.m > 0) || !(R.sp.dp_finest > 0)) return fail(c, f, "vrt_checkpoint_read: corrupt species record", VRT_ERR_ARG);
        R.d.resize(n);
        if (fread(R.d.data(), sizeof(vrt_patch_desc), n, f) != (size_t)n) return fail(c, f, "vrt_checkpoint_read: truncated file", VRT_ERR_ARG);
        off += (long)n * (long)sizeof(vrt_patch_desc);
        int path_now = 0;
        if (int rc = vrt_hierarchy_path(c, n, R.d.data(), &path_now)) { fclose(f); return rc; }
        if (path_now != R.path) return fail(c, f, "vrt_checkpoint_read: path mismatch (vrt_set_path)", VRT_ERR_STATE);
        long doubles = 0;
        if (R.path == VRT_PATH_FUSED) doubles = vrt_slab_plane_doubles(c->x_end - c->x_begin, R.d[0].n_p);
        else for (const vrt_patch_desc& q : R.d) doubles += (long)(q.n_x + 4) * (q.n_p + 4);
        R.data_off = off;
        off += doubles * (long)sizeof(double);
        if (off > file_size) return fail(c, f, "vrt_checkpoint_read: truncated file", VRT_ERR_ARG);
    }
    if (off != file_size) return fail(c, f, "vrt_checkpoint_read: trailing bytes (not a checkpoint of this configuration)", VRT_ERR_ARG);
    // ---- pass 2: apply ------------------------------------------------------------------------------------------------------
    Stager st(c, f);
    if (!st.ok()) return fail(c, f, "vrt_checkpoint_read: no pinned staging buffer", VRT_ERR_NOMEM);
    auto abort_restore = [&](const std::string& msg, int code) {
        vrt_invalidate_hierarchies(c);
        return fail(c, f, msg + " (the context holds no hierarchy now)", code);
    };
    if (fseek(f, (long)sizeof(Header), SEEK_SET)) return fail(c, f, "vrt_checkpoint_read: cannot seek", VRT_ERR_ARG);
    bool ok = true;
    for (int v = 0; v < 6 && ok; v++) ok = st.read(F.Y[v], 8L * F.M);
    ok = ok && st.read(F.a_squared, F.N + 1) && st.read(F.PHI, F.N) && st.read(F.neutral, F.N) && st.read(F.Ex0, 1);
    if (!ok) return abort_restore("vrt_checkpoint_read: read failed", VRT_ERR_CUDA);
    for (int s = 0; s < c->n_species; s++) {
        const Rec& R = recs[s];
        const int n = (int)R.d.size();
        int rc = vrt_set_species(c, s, R.sp.m, R.sp.q, R.sp.pmin, R.sp.dp_finest);
        if (!rc) rc = vrt_set_hierarchy(c, s, n, R.d.data());        // same storage as the writer's (zeroed)
        if (rc) { const std::string msg = c->err; return abort_restore(msg, rc); }
        VrtSpeciesState& S = c->S[s];
        if (fseek(f, R.data_off, SEEK_SET)) return abort_restore("vrt_checkpoint_read: cannot seek", VRT_ERR_ARG);
        if (S.path == VRT_PATH_FUSED) {
            ok = st.read(S.slab.f[0], (size_t)S.slab.plane);
            S.i_f0 = S.i_f1 = 0;
        } else {
            for (int p = 0; p < n && ok; p++) {
                ok = st.read(S.patches[p].f1, (size_t)S.patches[p].npad);
                if (ok && cudaMemcpyAsync(S.patches[p].f0, S.patches[p].f1, sizeof(double) * S.patches[p].npad, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) ok = false;
            }
        }
        if (!ok) return abort_restore("vrt_checkpoint_read: read failed", VRT_ERR_CUDA);
    }
    fclose(f);
    c->time = h.time;
    if (int rc = vrt_fields_refresh_efield(c)) return rc;     // the E table of the last Poisson solve, from PHI and Ex0
    VRT_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int vrt_checkpoint_read(vrt_ctx* c, const char* path) {
    if (!c || !path) return VRT_ERR_ARG;
    if (!c->grid_set) { c->err = "vrt_checkpoint_read: set the grid and the species first"; return VRT_ERR_STATE; }
    cudaSetDevice(c->device);
    try { return checkpoint_read_impl(c, path); }       // nothing throws across the C boundary
    catch (const std::exception&) { c->err = "vrt_checkpoint_read: out of host memory"; return VRT_ERR_NOMEM; }
}

}  // extern "C"
