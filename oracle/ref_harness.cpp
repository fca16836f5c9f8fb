// TEST INFRASTRUCTURE — oracle harness.  Not part of the product; nothing under veritas_b200/
// links or calls it.  It drives the UNMODIFIED reference classes (compiled from where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/) and dumps full-precision state so
// that the C restatement (oracle/veritas_oracle.c) and the CUDA path can be checked against the
// reference's own CPU solver.
//
// The reference is configured by editing its case file (docs/index.rst:114-160): this file plays
// the role of /root/reference/veritas.cpp — it supplies main() and the user-defined members of
// Settings (settingsOverride, RefinementOverride, GetBY, GetBZ, InitialDistribution), taking the
// laser-plasma case from veritas_b200/host/laser_plasma_case.hpp.  Private members of
// SolverManager / EMFieldSolver / Rectangle are read with g++ -fno-access-control.
//
// The same source also compiles against the veritas_b200 host classes (veritas_b200/host/, -DVRT_HOST_BUILD): that build
// (oracle/_ref/host_harness) is the drop-in test of the host layer — same case file, same dumps, GPU numerics.
//
// usage: ref_harness out.bin nx np Lfinest density steps [key=value ...]
//   keys: dump_every=1 stage_dumps=0 regrid_every=0 threads=N refine_mode=0 tail_p0=2 pre_steps=-1
//         time_only=0 internals=0 a0=1 file_output=0 (every N steps: SolverManager::fileOutput + OutputRectangles into
//         ./output, which must exist with its rectangleData subdirectory) precision=15 energy=0 (dN/dp dump; use threads=1)
//         compact=0 (1: per patch only state 1 of f, as record ".../f1" — at a step boundary states 0 and 1 coincide,
//         Rectangle.cpp:1614-1622 — a third of the dump size for the large per-step parity cases)
//         patch_moments=0 (1, reference build only: before every dump run EMFieldSolver::AssembleRhoAndJ on the dumped state and
//         add each patch's Rectangle::chargeR / currentR, what Level::CollectRhoAndJ sums, Level.cpp:42-62; the dumped charge /
//         J / charges are then the moments of that state, which the next Advance recomputes at its stage 0 anyway)
//         warmup=0 (time_only runs: that many steps run before the `steps` timed ones and are left out of ORACLE_TIMING)
//         ckpt_write=K (host build only: SolverManager::Checkpoint into <out>.ckpt after step K, with the driver loop's own
//         counters in <out>.ckpt.meta)  ckpt_restart=1 (host build only: skip the fields-only phase and the steps up to the
//         checkpoint, SolverManager::Restart from <out>.ckpt, continue with the remaining steps); $VRT_HARNESS_CKPT overrides the path
//
//        ref_harness cluster cases.txt out.txt nx np Lfinest
//   runs the regrid clustering members of Mesh (Mesh.cpp:298-792) on the flag sets of cases.txt, one case per line:
//     split <minEfficiency> <n> x p x p ...        getExtrema + splitRectangle      -> "k x0 p0 x1 p1 ..." (boxes, inclusive corners)
//     interp <lvl> <k> x0 p0 x1 p1 ...             interpRectanglesUp               -> "k x0 p0 x1 p1 ..."
//     merge <lvl> x0 p0 x1 p1                      mergeDownFlaggedData             -> "n x p x p ..."
//   The host build runs them without a Mesh and without a device (tests/test_host_clustering.py).
//
//        ref_harness settings out.txt nx np Lfinest [np_ion]
//   prints what Settings derives from Input / Particles (Settings.cpp:5-195): level sizes and spacings, species constants, fMax,
//   and the stage times of UpdateTime over two steps, as "name value" lines with 17 significant digits (no device needed).
//
//        ref_harness transfer out.txt
//   the host-side regrid data path on hand-made patches (no Mesh, no device): Rectangle::GetInterpolantsREF,
//   GetDataFromCoarseLevelRectangle / GetDataFromSameLevelRectangle / GetDataFromCoarseNewLevelRectangle (Rectangle.cpp:892-941,
//   1100-1128) from a coarse and an old fine patch into two new fine patches, and Rectangle::getError (866-890) on the coarse and
//   on a new patch, and Rectangle::InitializeDistribution (616-669) on a coarse and a fine patch; prints the patches' f and the
//   flagged cells with 17 significant digits.
#include "veritas.hpp"
#include "Settings.hpp"
#include "SolverManager.hpp"
#include "EMSolver.hpp"
#include "Mesh.hpp"
#include "Level.hpp"
#include "Rectangle.hpp"
#include "BoundaryCondition.hpp"
#include "../veritas_b200/host/laser_plasma_case.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

bool LOUD = false, NOISY = false;
static vrt_case::LaserPlasma g_case;

void Settings::settingsOverride() {
    unsigned ps[2] = {p_size[0], p_size[1]};
    vrt_case::Derived d = vrt_case::derive(g_case, m[0], q[0], x_size, ps, refinementCriteria);
    dp[0] = d.dp[0]; dp[1] = d.dp[1];
    pmin[0] = d.pmin[0]; pmin[1] = d.pmin[1];
    dx = d.dx;
    sizeWeight = d.sizeWeight;
    quadratureDepth = d.quadratureDepth;
    temp[0].resize(2, 0.0); temp[0][0] = d.temp0[0]; temp[0][1] = d.temp1[0];
    temp[1].resize(2, 0.0); temp[1][0] = d.temp0[1]; temp[1][1] = d.temp1[1];
    tempEM.resize(2, 0.0); tempEM[0] = d.tempEM[0]; tempEM[1] = d.tempEM[1];
}
bool Settings::RefinementOverride(double x, double p, int depth, int particleType) {
    return vrt_case::refine_override(g_case, x, p, temp[particleType][1], m[0]);
}
double Settings::GetBY(double x, double t) { return vrt_case::laser_by(tempEM[0], tempEM[1], x, t); }
double Settings::GetBZ(double x, double t) { return vrt_case::laser_bz(tempEM[0], tempEM[1], x, t); }
double Settings::InitialDistribution(double x, double p, int particleType) {
    return vrt_case::maxwellian_slab(x, p, plasma_xl_bound, plasma_xr_bound, temp[particleType][0], temp[particleType][1]);
}

// ---- tiny binary container: records of  name | ndim | dims | float64 data -------------------
static FILE* g_out = nullptr;
static void put(const std::string& name, const double* data, std::initializer_list<long> dims) {
    unsigned nl = name.size(), nd = dims.size();
    fwrite(&nl, 4, 1, g_out); fwrite(name.data(), 1, nl, g_out); fwrite(&nd, 4, 1, g_out);
    long n = 1;
    for (long d : dims) { fwrite(&d, 8, 1, g_out); n *= d; }
    fwrite(data, 8, n, g_out);
}
static void put1(const std::string& name, double v) { put(name, &v, {1}); }

static bool g_compact = false, g_patch_moments = false;
static void dump_state(SolverManager& SM, Settings& st, const std::string& tag, bool internals) {
#ifdef VRT_HOST_BUILD
    SM.SyncHost();     // veritas_b200 host classes: Rectangle::f and the EMFieldSolver arrays are mirrors of device data
#else
    if (g_patch_moments) SM.EMSolver->AssembleRhoAndJ();
#endif
    EMFieldSolver& em = *SM.EMSolver;
    long M = em.x_size + em.n_prepad + em.n_postpad, N = em.x_size;
    put1(tag + "/time", st.time);
    put(tag + "/By", em.By.data(), {8, M}); put(tag + "/Bz", em.Bz.data(), {8, M});
    put(tag + "/Ey", em.Ey.data(), {8, M}); put(tag + "/Ez", em.Ez.data(), {8, M});
    put(tag + "/Ay", em.Ay.data(), {8, M}); put(tag + "/Az", em.Az.data(), {8, M});
    put(tag + "/a_squared", em.a_squared.data(), {N + 1});
    put(tag + "/PHI", em.PHI, {N});
    put1(tag + "/Ex0", em.Ex0);
    put(tag + "/charge", em.charge.data(), {N});
    put(tag + "/J", em.J.data(), {N});
    put(tag + "/neutralizationCharge", em.neutralizationCharge.data(), {N});
    for (size_t s = 0; s < SM.meshes.size(); s++) {
        put(tag + "/charges" + std::to_string(s), em.charges[s].data(), {N});
        Mesh& mesh = *SM.meshes[s];
        for (size_t l = 0; l < mesh.levels.size(); l++) {
            auto& rects = mesh.levels[l]->rectangles;
            for (size_t r = 0; r < rects.size(); r++) {
                Rectangle& R = *rects[r];
                std::string p = tag + "/s" + std::to_string(s) + "/l" + std::to_string(l) + "/r" + std::to_string(r);
                double desc[10] = {(double)R.n_x, (double)R.n_p, (double)R.x_pos, (double)R.p_pos, (double)R.depth,
                                   (double)R.up, (double)R.down, (double)R.left, (double)R.right, R.relativeToBottom};
                put(p + "/desc", desc, {10});
                if (g_compact) {
                    std::vector<double> f1((size_t)(R.n_x + 4) * (R.n_p + 4));
                    for (size_t c = 0; c < f1.size(); c++) f1[c] = R.f[3 * c + 1];
                    put(p + "/f1", f1.data(), {R.n_x + 4, R.n_p + 4});
                } else {
                    put(p + "/f", R.f.data(), {R.n_x + 4, R.n_p + 4, 3});
                }
#ifndef VRT_HOST_BUILD
                if (g_patch_moments) {
                    put(p + "/chargeR", R.chargeR.data(), {(long)R.chargeR.size()});
                    put(p + "/currentR", R.currentR.data(), {(long)R.currentR.size()});
                }
                if (internals) {
                    put(p + "/FxH", R.FxH.data(), {R.n_x + 4, R.n_p + 4, 6});
                    put(p + "/FpH", R.FpH.data(), {R.n_x + 4, R.n_p + 4, 6});
                    put(p + "/FxL", R.FxL.data(), {R.n_x + 4, R.n_p + 4, 6});
                    put(p + "/FpL", R.FpL.data(), {R.n_x + 4, R.n_p + 4, 6});
                    put(p + "/FxDS", R.FxDS.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/FpDS", R.FpDS.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/Rp", R.Rp.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/Rm", R.Rm.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/Cx", R.Cx.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/Cp", R.Cp.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/ex", R.ex.data(), {R.n_x + 4, R.n_p + 4});
                    put(p + "/ep", R.ep.data(), {R.n_x + 4, R.n_p + 4});
                }
#endif
            }
        }
    }
}

// the laser-plasma case's Input / Particles (veritas.cpp:118-133 with the sizes as parameters)
static void configure_case(unsigned nx, unsigned np, unsigned np_ion, unsigned Lfinest, Input& grid, Particles& particles) {
    grid.minEfficiency = 0.75; grid.dx = 0.5; grid.k = 0.01;
    grid.refinementCriteria = 1e-8; grid.cfl = 0.5; grid.sizeWeight = 0.0;
    grid.preLength = 0; grid.postLength = 0;
    grid.nx = nx; grid.r = 2; grid.Lfinest = Lfinest;
    grid.tempEM = {0.0};
    particles.mass = {9.10938291e-31, 9.10938291e-31 * 1836};
    particles.charge = {-1.60217657e-19, 1.60217657e-19};
    particles.misc = {{0.0, 0.01}, {0.0, 0.01}};
    particles.np = {np, np_ion};
    particles.dp = {0.1, 0.1};
    particles.pmin = {0.1, 0.1};
}

static void put_boxes(std::ostream& o, const level& L) {
    o << L.size();
    for (const rect& r : L) o << ' ' << r.first.first << ' ' << r.first.second << ' ' << r.second.first << ' ' << r.second.second;
    o << '\n';
}

// clustering members on given flag sets (see the usage comment)
static int cluster_mode(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s cluster cases.txt out.txt nx np Lfinest\n", argv[0]); return 2; }
    std::ifstream in(argv[2]);
    std::ofstream out(argv[3]);
    if (!in || !out) { fprintf(stderr, "cluster: cannot open %s / %s\n", argv[2], argv[3]); return 1; }
    g_case.density = 0.1;
    Input grid; Particles particles; Output output;
    configure_case(atoi(argv[4]), atoi(argv[5]), atoi(argv[5]), atoi(argv[6]), grid, particles);
    output.time = output.rectangleData = output.charge = output.potential = output.EFieldLongitudinal =
        output.EFieldTransverse = output.BFieldTransverse = output.AFieldSquared = output.energy = false;
    Settings settings(grid, particles, output);
#ifndef VRT_HOST_BUILD
    SolverManager SM(settings);
    Mesh& M = *SM.meshes[0];
#else
    // level sizes as Mesh::interpRectanglesUp / mergeDownFlaggedData take them from Settings (species 0)
    auto nx_of = [&](int lvl) { return settings.GetXSize(settings.maxDepth - lvl); };
    auto np_of = [&](int lvl) { return settings.GetPSize(settings.maxDepth - lvl, 0); };
#endif
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string kind;
        if (!(ls >> kind)) continue;
        if (kind == "split") {
            double eff; int n;
            ls >> eff >> n;
            std::vector<coords> flagged(n);
            for (coords& c : flagged) ls >> c.first >> c.second;
            rect ext;
#ifndef VRT_HOST_BUILD
            M.getExtrema(ext, flagged);
            put_boxes(out, M.splitRectangle(ext, flagged, eff));
#else
            Mesh::getExtrema(ext, flagged);
            put_boxes(out, Mesh::splitRectangle(ext, flagged, eff));
#endif
        } else if (kind == "interp") {
            int lvl, k;
            ls >> lvl >> k;
            level L(k);
            for (rect& r : L) ls >> r.first.first >> r.first.second >> r.second.first >> r.second.second;
#ifndef VRT_HOST_BUILD
            M.interpRectanglesUp(L, lvl);
#else
            Mesh::scaleRectanglesUp(L, nx_of(lvl), np_of(lvl), nx_of(lvl + 1), np_of(lvl + 1));
#endif
            put_boxes(out, L);
        } else if (kind == "merge") {
            int lvl;
            rect r;
            ls >> lvl >> r.first.first >> r.first.second >> r.second.first >> r.second.second;
            std::vector<coords> found;
#ifndef VRT_HOST_BUILD
            M.mergeDownFlaggedData(lvl, r, found);
#else
            Mesh::footprintBelow(r, nx_of(lvl), np_of(lvl), nx_of(lvl + 2), np_of(lvl + 2), (int)settings.refinementRatio, found);
#endif
            out << found.size();
            for (const coords& c : found) out << ' ' << c.first << ' ' << c.second;
            out << '\n';
        } else { fprintf(stderr, "cluster: unknown case kind %s\n", kind.c_str()); return 2; }
    }
    return 0;
}

// Settings' derived quantities (see the usage comment)
static int settings_mode(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s settings out.txt nx np Lfinest [np_ion]\n", argv[0]); return 2; }
    FILE* o = fopen(argv[2], "w");
    if (!o) { perror(argv[2]); return 1; }
    g_case.density = 0.1;
    Input grid; Particles particles; Output output;
    configure_case(atoi(argv[3]), atoi(argv[4]), argc > 6 ? atoi(argv[6]) : atoi(argv[4]), atoi(argv[5]), grid, particles);
    output.time = output.rectangleData = output.charge = output.potential = output.EFieldLongitudinal =
        output.EFieldTransverse = output.BFieldTransverse = output.AFieldSquared = output.energy = false;
    Settings st(grid, particles, output);
    auto P = [&](const std::string& name, double v) { fprintf(o, "%s %.17g\n", name.c_str(), v); };
    P("maxDepth", st.maxDepth); P("refinementRatio", st.refinementRatio); P("x_size", st.x_size); P("x_size_finest", st.x_size_finest);
    P("dx", st.dx); P("minEfficiency", st.minEfficiency); P("refinementCriteria", st.refinementCriteria); P("cfl", st.cfl);
    P("sizeWeight", st.sizeWeight); P("quadratureDepth", st.quadratureDepth); P("preLength", st.preLength); P("postLength", st.postLength);
    P("plasma_xl_bound", st.plasma_xl_bound); P("plasma_xr_bound", st.plasma_xr_bound);
    P("tempEM0", st.tempEM[0]); P("tempEM1", st.tempEM[1]);
    for (int l = 0; l <= st.maxDepth; l++) {
        P("GetDx" + std::to_string(l), st.GetDx(l)); P("GetXSize" + std::to_string(l), st.GetXSize(l));
        for (int s = 0; s < 2; s++) {
            P("GetDp" + std::to_string(l) + "_" + std::to_string(s), st.GetDp(l, s));
            P("GetPSize" + std::to_string(l) + "_" + std::to_string(s), st.GetPSize(l, s));
        }
    }
    for (int s = 0; s < 2; s++) {
        const std::string k = std::to_string(s);
        P("m" + k, st.GetMass(s)); P("q" + k, st.GetCharge(s)); P("dp" + k, st.dp[s]); P("pmin" + k, st.pmin[s]);
        P("p_size" + k, st.p_size[s]); P("p_size_finest" + k, st.p_size_finest[s]); P("temp0_" + k, st.temp[s][0]); P("temp1_" + k, st.temp[s][1]);
        P("fMax" + k, st.GetfMax(s));
    }
    const double dts[2] = {1.25e-18, 3.0e-17};
    for (int n = 0; n < 2; n++)
        for (int i = 0; i < 6; i++) { st.UpdateTime(i, dts[n]); P("time_step" + std::to_string(n) + "_stage" + std::to_string(i), st.time); }
    fclose(o);
    return 0;
}

// host-side regrid data path on hand-made patches (see the usage comment)
static int transfer_mode(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s transfer out.txt\n", argv[0]); return 2; }
    FILE* o = fopen(argv[2], "w");
    if (!o) { perror(argv[2]); return 1; }
    g_case.density = 0.1;
    Input grid; Particles particles; Output output;
    configure_case(32, 16, 16, 2, grid, particles);
    output.time = output.rectangleData = output.charge = output.potential = output.EFieldLongitudinal =
        output.EFieldTransverse = output.BFieldTransverse = output.AFieldSquared = output.energy = false;
    Settings st(grid, particles, output);
    st.refinementCriteria = 0.02;            // of the order of the test pattern's scaled differences: a non-trivial flag set
    std::shared_ptr<Rectangle> bc = std::make_shared<BoundaryCondition>();
    auto make = [&](int n_x, int n_p, int x_pos, int p_pos, int depth) {
        return std::make_shared<Rectangle>(n_x, n_p, x_pos, p_pos, depth, st, bc, true, true, true, true, 0);
    };
    const double fm = st.GetfMax(0);
    // smooth bump + ripple + floor in finest-grid units, different in states 0 and 1, ghost layers included
    auto fill = [&](Rectangle& R, double shift) {
        for (int i = -2; i < R.n_x + 2; i++)
            for (int j = -2; j < R.n_p + 2; j++) {
                const double X = (i + R.x_pos + 0.5) * R.relativeToBottom, P = (j + R.p_pos + 0.5) * R.relativeToBottom;
                const double v = fm * (std::exp(-((X - 30 - shift) * (X - 30 - shift) / 80 + (P - 14) * (P - 14) / 20)) +
                                       0.05 * std::sin(0.7 * X) * std::cos(0.9 * P) + 0.06);
                R.f[R.Index3(i, j, 0)] = v;
                R.f[R.Index3(i, j, 1)] = v * (1 + 0.01 * std::cos(X + shift));
                R.f[R.Index3(i, j, 2)] = -v;
            }
    };
    auto dump_f = [&](const char* name, Rectangle& R) {
        fprintf(o, "%s %d %d %d %d\n", name, R.n_x, R.n_p, R.x_pos, R.p_pos);
        for (int i = -2; i < R.n_x + 2; i++)
            for (int j = -2; j < R.n_p + 2; j++)
                fprintf(o, "%d %d %.17g %.17g %.17g\n", i, j, R.f[R.Index3(i, j, 0)], R.f[R.Index3(i, j, 1)], R.f[R.Index3(i, j, 2)]);
    };
    auto dump_flags = [&](const char* name, Rectangle& R) {
        std::vector<coords> flagged;
        R.getError(flagged, 0);
        fprintf(o, "%s %zu", name, flagged.size());
        for (const coords& c : flagged) fprintf(o, " %d %d", c.first, c.second);
        fprintf(o, "\n");
    };
    auto coarse = make(32, 16, 0, 0, 1);          // the whole coarse level
    auto old_fine = make(16, 8, 10, 6, 0);        // fine-level indices (64 x 32 grid)
    auto new_a = make(24, 12, 8, 4, 0);           // overlaps old_fine: coarse interpolation, then same-level copy on the overlap
    auto new_b = make(8, 8, 40, 16, 0);           // no old fine data: filled by the coarse-new pass alone
    auto new_c = make(12, 8, 0, 24, 0);           // touches the domain corner: the coarse ring reaches into the ghost layers
    fill(*coarse, 0.0); fill(*old_fine, 1.5);
    const double five[3][5] = {{0.1, 0.4, 0.9, 0.5, 0.2}, {1.0, 1.0, 1.0, 1.0, 1.0}, {0.0, -0.3, 2.0, 0.7, 0.1}};
    for (const double* v : {five[0], five[1], five[2]}) {
        std::vector<double> sub = coarse->GetInterpolantsREF(v[0], v[1], v[2], v[3], v[4]);
        fprintf(o, "interpolants %zu", sub.size());
        for (double x : sub) fprintf(o, " %.17g", x);
        fprintf(o, "\n");
    }
    // the order of Mesh::InterMeshDataTransfer (Mesh.cpp:116-130)
    for (auto& target : {new_a, new_b, new_c}) coarse->GetDataFromCoarseLevelRectangle(target);
    dump_f("after_coarse_a", *new_a);
    for (auto& target : {new_a, new_b, new_c}) old_fine->GetDataFromSameLevelRectangle(target);
    dump_f("after_same_a", *new_a);
    auto fresh_b = make(8, 8, 40, 16, 0);
    auto fresh_a = make(24, 12, 8, 4, 0);
    old_fine->GetDataFromSameLevelRectangle(fresh_a);
    for (auto& target : {fresh_a, fresh_b}) coarse->GetDataFromCoarseNewLevelRectangle(target);
    dump_f("coarse_new_a", *fresh_a); dump_f("coarse_new_b", *fresh_b); dump_f("corner_c", *new_c);
    dump_flags("flags_coarse", *coarse); dump_flags("flags_new_a", *new_a); dump_flags("flags_old_fine", *old_fine);
    // Rectangle::InitializeDistribution (Rectangle.cpp:616-669): sub-cell quadrature of the user's initial distribution, on a coarse
    // patch and on a fine patch across the plasma slab's left edge
    auto init_coarse = make(32, 16, 0, 0, 1);
    auto init_fine = make(16, 8, 14, 12, 0);
    init_coarse->InitializeDistribution(); init_fine->InitializeDistribution();
    dump_f("init_coarse", *init_coarse); dump_f("init_fine", *init_fine);
    fclose(o);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "transfer") return transfer_mode(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "cluster") return cluster_mode(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "settings") return settings_mode(argc, argv);
    if (argc < 7) { fprintf(stderr, "usage: %s out.bin nx np Lfinest density steps [key=value...]\n", argv[0]); return 2; }
    const char* out = argv[1];
    unsigned nx = atoi(argv[2]), np = atoi(argv[3]), Lfinest = atoi(argv[4]);
    g_case.density = atof(argv[5]);
    int steps = atoi(argv[6]);
    std::map<std::string, double> kv = {{"dump_every", 1}, {"stage_dumps", 0}, {"regrid_every", 0}, {"threads", 0},
                                        {"refine_mode", 0}, {"tail_p0", 2}, {"pre_steps", -1}, {"time_only", 0},
                                        {"internals", 0}, {"a0", 1}, {"np_ion", 0},
                                        {"file_output", 0}, {"precision", 15}, {"energy", 0}, {"compact", 0}, {"patch_moments", 0}, {"warmup", 0}, {"ckpt_write", 0}, {"ckpt_restart", 0}};
    for (int i = 7; i < argc; i++) {
        std::string a = argv[i]; size_t e = a.find('=');
        if (e == std::string::npos || !kv.count(a.substr(0, e))) { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
        kv[a.substr(0, e)] = atof(a.c_str() + e + 1);
    }
    g_case.refine_mode = (int)kv["refine_mode"]; g_case.tail_p0 = kv["tail_p0"]; g_case.a0 = kv["a0"];
    if (kv["threads"] > 0) omp_set_num_threads((int)kv["threads"]);
    bool time_only = kv["time_only"] != 0, internals = kv["internals"] != 0;
    g_compact = kv["compact"] != 0; g_patch_moments = kv["patch_moments"] != 0;

    Input grid; Particles particles; Output output;
    unsigned np_ion = kv["np_ion"] > 0 ? (unsigned)kv["np_ion"] : np;
    configure_case(nx, np, np_ion, Lfinest, grid, particles);
    const int file_output = (int)kv["file_output"];
    output.precision = (int)kv["precision"];
    // dN/dp: the reference accumulates it with a data race — only deterministic with threads=1
    output.energy = kv["energy"] != 0;
    if (!file_output) output.time = output.rectangleData = output.charge = output.potential = output.EFieldLongitudinal =
        output.EFieldTransverse = output.BFieldTransverse = output.AFieldSquared = false;

    Settings settings(grid, particles, output);
    SolverManager SM(settings);

    if (!time_only) {
        g_out = fopen(out, "wb");
        if (!g_out) { perror(out); return 1; }
    }
    double T = settings.tempEM[0] / cs;
    double dt = T / 400, t = dt;
    int pre_steps = (int)kv["pre_steps"];
    int n_pre = 0;
    const int ckpt_write = (int)kv["ckpt_write"];
    const bool ckpt_restart = kv["ckpt_restart"] != 0;
    const std::string ckpt_path = getenv("VRT_HARNESS_CKPT") ? std::string(getenv("VRT_HARNESS_CKPT")) : std::string(out) + ".ckpt";
    int first_step = 1, counter = 0;
#ifndef VRT_HOST_BUILD
    if (ckpt_write || ckpt_restart) { fprintf(stderr, "ckpt_write / ckpt_restart need the veritas_b200 host build\n"); return 2; }
#else
    if (ckpt_restart) {
        SM.Restart(ckpt_path);
        FILE* m = fopen((ckpt_path + ".meta").c_str(), "r");
        if (!m || fscanf(m, "%d %d %d %la", &first_step, &counter, &n_pre, &t) != 4) { fprintf(stderr, "cannot read %s.meta\n", ckpt_path.c_str()); return 1; }
        fclose(m);
        first_step += 1;
    }
#endif
    // fields-only phase, exactly as the shipped main(): while t <= 3T (veritas.cpp:135-144)
    while (!ckpt_restart && (pre_steps < 0 ? !(t > 3 * T) : n_pre < pre_steps)) {
        SM.AdvanceFields(dt);
        t += dt; n_pre++;
    }
    if (!time_only) {
        double meta[8] = {(double)nx, (double)np, (double)Lfinest, g_case.density, (double)steps, (double)n_pre, dt, t};
        put("meta", meta, {8});
        put1("dx", settings.GetDx(0));
        for (int s = 0; s < 2; s++) {
            double sp[4] = {settings.m[s], settings.q[s], settings.pmin[s], settings.GetDp(0, s)};
            put("species" + std::to_string(s), sp, {4});
        }
        double em[2] = {settings.tempEM[0], settings.tempEM[1]};
        put("tempEM", em, {2});
        if (!ckpt_restart) dump_state(SM, settings, "step0", internals);
    }
    int regrid_every = (int)kv["regrid_every"], dump_every = (int)kv["dump_every"];
    bool stage_dumps = kv["stage_dumps"] != 0;
    // timing (CPU baseline / drop-in comparison): CalculateDt + Advance of every step, as in the reference's main loop
    // (veritas.cpp:137-144).  The GPU build's Advance only enqueues work; the next CalculateDt reads the CFL bound back and
    // thereby waits for it, so the last step is closed by one more CalculateDt inside the timed region.
    double advance_seconds = 0.0, cell_updates = 0.0;
    auto count_cells = [&]() {
        long c = 0;
        for (auto& mesh : SM.meshes) for (auto& lvl : mesh->levels) for (auto& r : lvl->rectangles) c += (long)r->n_x * r->n_p;
        return c;
    };
    long cells = count_cells();
    const int warmup = time_only ? (int)kv["warmup"] : 0;
    long launches = 0;
    steps += warmup;
    for (int n = first_step; n <= steps; n++) {
        auto t0 = std::chrono::steady_clock::now();
        double dt_adaptive = std::min(SM.CalculateDt(settings.cfl), dt);
        if (stage_dumps && !time_only) {
            // SolverManager::Advance unrolled (SolverManager.cpp:28-39) with a dump after every stage
            for (int i = 0; i < 6; i++) {
                SM.EMSolver->AssembleRhoAndJ();
                SM.EMSolver->UpdatePotential();
                for (unsigned j = 0; j < settings.q.size(); j++) SM.meshes[j]->Advance(dt_adaptive, i);
                settings.UpdateTime(i, dt_adaptive);
                SM.EMSolver->RGKStep(i, dt_adaptive);
                dump_state(SM, settings, "step" + std::to_string(n) + "_stage" + std::to_string(i), internals);
            }
        } else {
            SM.Advance(dt_adaptive);
        }
        if (time_only && (n == steps || n == warmup || (regrid_every > 0 && counter >= regrid_every - 1))) SM.CalculateDt(settings.cfl);
        if (n > warmup) {
            advance_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            cell_updates += 6.0 * (double)cells;
#ifdef VRT_HOST_BUILD
            launches += vrt_last_step_launches(settings.Gpu());
#endif
        }
        if (regrid_every > 0) {
            if (counter >= regrid_every - 1) { SM.reGrid(t); counter = 0; cells = count_cells(); } else counter++;
        }
        t += dt_adaptive;
        if (file_output > 0 && n % file_output == 0) { SM.fileOutput(t); SM.OutputRectangles(t); }
        if (!time_only && (n % dump_every == 0 || n == steps)) {
            put1("step" + std::to_string(n) + "/dt", dt_adaptive);
            dump_state(SM, settings, "step" + std::to_string(n), internals);
        }
#ifdef VRT_HOST_BUILD
        if (ckpt_write > 0 && n == ckpt_write) {
            SM.Checkpoint(ckpt_path);
            FILE* m = fopen((ckpt_path + ".meta").c_str(), "w");
            fprintf(m, "%d %d %d %a\n", n, counter, n_pre, t);
            fclose(m);
        }
#endif
    }
    if (g_out) fclose(g_out);
    // machine-readable timing line (CPU baseline): cells, steps, seconds in Advance, threads
    printf("ORACLE_TIMING cells=%ld steps=%d advance_s=%.6f threads=%d cell_updates_per_s_per_stage=%.6e gpu_launches=%ld\n", cells, steps - warmup,
           advance_seconds, omp_get_max_threads(), steps > warmup ? cell_updates / advance_seconds : 0.0, launches);
    return 0;
}
