// TEST INFRASTRUCTURE.  The drop-in boundary as a compile check: every member of the reference's public C++
// class surface on the hot path — the one /root/reference/veritas.cpp and oracle/ref_harness.cpp program against — named with
// its exact type.  The SAME file must compile (g++ -fsyntax-only) against the unmodified reference headers and against the
// veritas_b200 host classes (-DVRT_HOST_BUILD -Iveritas_b200/host); tests/test_host_api_surface.py does both.
#include "veritas.hpp"
#include "Settings.hpp"
#include "SolverManager.hpp"
#include "EMSolver.hpp"
#include "Mesh.hpp"
#include "Level.hpp"
#include "Rectangle.hpp"
#include "BoundaryCondition.hpp"
#include <memory>
#include <type_traits>
#include <vector>

bool LOUD = false, NOISY = false;

template <class T> static void use(T) {}
#define MEMBER(Class, name, ...) use(static_cast<__VA_ARGS__>(&Class::name))
#define FIELD(Class, name, Type) static_assert(std::is_same<decltype(Class::name), Type>::value, #Class "::" #name)

void surface() {
    // Input / Particles / Output and Settings (Settings.hpp:4-64)
    Input in; Particles pa; Output ou;
    in.nx = 8; in.r = 2; in.Lfinest = 1; in.dx = 0.5; in.minEfficiency = 0.75; in.refinementCriteria = 1e-8; in.cfl = 0.5;
    in.sizeWeight = 0.0; in.preLength = 0; in.postLength = 0; in.k = 0.01; in.tempEM = {0.0};
    in.plasma_xl_bound = 3e-6; in.plasma_xr_bound = 7e-6;
    pa.mass = {1.0}; pa.charge = {-1.0}; pa.misc = {{0.0, 0.01}}; pa.np = {8u}; pa.dp = {0.1}; pa.pmin = {0.1};
    ou.precision = 15;
    ou.time = ou.rectangleData = ou.charge = ou.potential = ou.EFieldLongitudinal = ou.EFieldTransverse = ou.BFieldTransverse =
        ou.AFieldSquared = ou.energy = false;
    use(static_cast<Settings* (*)(const Input&, const Particles&, const Output&)>(
        [](const Input& a, const Particles& b, const Output& c) { return new Settings(a, b, c); }));
    FIELD(Settings, output, Output); FIELD(Settings, dx, double); FIELD(Settings, time, double); FIELD(Settings, cfl, double);
    FIELD(Settings, minEfficiency, double); FIELD(Settings, refinementCriteria, double); FIELD(Settings, sizeWeight, double);
    FIELD(Settings, plasma_xl_bound, double); FIELD(Settings, plasma_xr_bound, double);
    FIELD(Settings, m, std::vector<double>); FIELD(Settings, q, std::vector<double>); FIELD(Settings, dp, std::vector<double>);
    FIELD(Settings, pmin, std::vector<double>); FIELD(Settings, tempEM, std::vector<double>); FIELD(Settings, fMax, std::vector<double>);
    FIELD(Settings, temp, std::vector<std::vector<double>>);
    FIELD(Settings, x_size, unsigned int); FIELD(Settings, x_size_finest, unsigned int); FIELD(Settings, refinementRatio, unsigned int);
    FIELD(Settings, maxDepth, int); FIELD(Settings, quadratureDepth, int);
    FIELD(Settings, p_size, std::vector<unsigned int>); FIELD(Settings, p_size_finest, std::vector<unsigned int>);
    MEMBER(Settings, settingsOverride, void (Settings::*)());
    MEMBER(Settings, RefinementOverride, bool (Settings::*)(double, double, int, int));
    MEMBER(Settings, GetBY, double (Settings::*)(double, double));
    MEMBER(Settings, GetBZ, double (Settings::*)(double, double));
    MEMBER(Settings, InitialDistribution, double (Settings::*)(double, double, int));
    MEMBER(Settings, GetDp, double (Settings::*)(int, int));
    MEMBER(Settings, GetDx, double (Settings::*)(int));
    MEMBER(Settings, GetXSize, int (Settings::*)(int));
    MEMBER(Settings, GetPSize, int (Settings::*)(int, int));
    MEMBER(Settings, GetMass, double (Settings::*)(int));
    MEMBER(Settings, GetCharge, double (Settings::*)(int));
    MEMBER(Settings, GetfMax, double (Settings::*)(int));
    MEMBER(Settings, UpdateTime, void (Settings::*)(int, double));
    MEMBER(Settings, DetermineMaximum, void (Settings::*)());

    // SolverManager (SolverManager.hpp:5-19)
    use(static_cast<SolverManager* (*)(Settings&)>([](Settings& s) { return new SolverManager(s); }));
    MEMBER(SolverManager, Advance, void (SolverManager::*)(double));
    MEMBER(SolverManager, AdvanceFields, void (SolverManager::*)(double));
    MEMBER(SolverManager, reGrid, void (SolverManager::*)(double));
    MEMBER(SolverManager, CalculateDt, double (SolverManager::*)(double));
    MEMBER(SolverManager, fileOutput, void (SolverManager::*)(double));
    MEMBER(SolverManager, OutputRectangles, void (SolverManager::*)(double));

    // Mesh (Mesh.hpp:4-41)
    use(static_cast<Mesh* (*)(int, Settings&)>([](int t, Settings& s) { return new Mesh(t, s); }));
    FIELD(Mesh, particleType, int); FIELD(Mesh, levels, std::vector<std::unique_ptr<Level>>);
    MEMBER(Mesh, Advance, void (Mesh::*)(double, int));
    MEMBER(Mesh, PushData, void (Mesh::*)(int));
    MEMBER(Mesh, PushBoundaryC, void (Mesh::*)());
    MEMBER(Mesh, updateHierarchy, void (Mesh::*)(bool));
    MEMBER(Mesh, InterpolateRhoAndJToFinestMesh, void (Mesh::*)(std::vector<double>&, std::vector<double>&));
    MEMBER(Mesh, InterpolateEnergyToFinestMesh, void (Mesh::*)(std::vector<double>&));
    MEMBER(Mesh, SetFieldSolver, void (Mesh::*)(const std::shared_ptr<EMFieldSolver>&));
    MEMBER(Mesh, outputRectangleData, void (Mesh::*)(double));
    MEMBER(Mesh, promoteHierarchyToMesh, void (Mesh::*)(bool));
    MEMBER(Mesh, InterMeshDataTransfer, void (Mesh::*)(const std::vector<std::unique_ptr<Level>>&));
    MEMBER(Mesh, getError, void (Mesh::*)(const int&, bool, std::vector<coords>&));
    MEMBER(Mesh, interpRectanglesUp, void (Mesh::*)(level&, const int&));
    MEMBER(Mesh, mergeDownFlaggedData, void (Mesh::*)(const int&, const rect&, std::vector<coords>&));
    // default arguments and the clustering calls as a caller writes them (member or static: both must accept this syntax)
    use(static_cast<void (*)(Mesh&)>([](Mesh& m) {
        m.PushData(); m.updateHierarchy();
        std::vector<coords> flagged{{1, 1}};
        rect box; m.getExtrema(box, flagged);
        auto sig = m.computeSignatures(box, flagged);
        coords cut = m.identifyInflection(box, sig); (void)cut;
        std::tuple<bool, int, int> hole = m.hasHole(std::get<0>(sig)); (void)hole;
        level found = m.splitRectangle(box, flagged, 0.75);
        int n = m.countCells(box); (void)n; (void)found;
    }));

    // Level (Level.hpp:4-23)
    use(static_cast<Level* (*)(int, int, Settings&)>([](int t, int d, Settings& s) { return new Level(t, d, s); }));
    FIELD(Level, rectangles, std::vector<std::shared_ptr<Rectangle>>);
    MEMBER(Level, FCTTimeStep, void (Level::*)(double, int, int));
    MEMBER(Level, PushData, void (Level::*)(int, int));
    MEMBER(Level, CollectRhoAndJ, void (Level::*)());
    MEMBER(Level, CollectEnergy, void (Level::*)());
    MEMBER(Level, InterpolateRhoAndJToFinestMesh, void (Level::*)(std::vector<double>&, std::vector<double>&));
    MEMBER(Level, InterpolateEnergyToFinestMesh, void (Level::*)(std::vector<double>&));
    MEMBER(Level, GetDataFromSameLevel, void (Level::*)(const std::unique_ptr<Level>&));
    MEMBER(Level, GetDataFromCoarserLevel, void (Level::*)(const std::unique_ptr<Level>&));
    MEMBER(Level, GetDataFromCoarseNewLevel, void (Level::*)(const std::unique_ptr<Level>&));

    // Rectangle / BoundaryCondition (Rectangle.hpp:7-80, BoundaryCondition.hpp:7-11)
    use(static_cast<Rectangle* (*)(Settings&, const std::shared_ptr<Rectangle>&)>([](Settings& s, const std::shared_ptr<Rectangle>& bc) {
        return new Rectangle(8, 8, 0, 0, 0, s, bc, true, true, true, true, 0);
    }));
    FIELD(Rectangle, n_x, int); FIELD(Rectangle, n_p, int); FIELD(Rectangle, x_pos, int); FIELD(Rectangle, p_pos, int);
    FIELD(Rectangle, relativeToBottom, double);
    FIELD(Rectangle, f, std::vector<double>); FIELD(Rectangle, chargeR, std::vector<double>);
    FIELD(Rectangle, energyR, std::vector<double>); FIELD(Rectangle, currentR, std::vector<double>);
    MEMBER(Rectangle, GetValueFromSameLevel, double (Rectangle::*)(int, int, int));
    MEMBER(Rectangle, InitializeDistribution, void (Rectangle::*)());
    MEMBER(Rectangle, getError, void (Rectangle::*)(std::vector<coords>&, int));
    MEMBER(Rectangle, GetInterpolantsREF, std::vector<double> (Rectangle::*)(double, double, double, double, double));
    MEMBER(Rectangle, GetWenoValueFromCoarseLevel, std::vector<double> (Rectangle::*)(int, int, int, int));
    MEMBER(Rectangle, GetDataFromCoarseLevelRectangle, void (Rectangle::*)(const std::shared_ptr<Rectangle>&));
    MEMBER(Rectangle, GetDataFromSameLevelRectangle, void (Rectangle::*)(const std::shared_ptr<Rectangle>&));
    MEMBER(Rectangle, GetDataFromCoarseNewLevelRectangle, void (Rectangle::*)(const std::shared_ptr<Rectangle>&));
    static_assert(std::is_base_of<Rectangle, BoundaryCondition>::value, "BoundaryCondition : Rectangle");
    static_assert(std::has_virtual_destructor<Rectangle>::value || std::is_polymorphic<Rectangle>::value, "Rectangle is polymorphic");
    MEMBER(BoundaryCondition, GetValueFromSameLevel, double (BoundaryCondition::*)(int, int, int));

    // EMFieldSolver (EMSolver.hpp:9-63)
    use(static_cast<EMFieldSolver* (*)(Settings&, const std::vector<std::shared_ptr<Mesh>>&)>(
        [](Settings& s, const std::vector<std::shared_ptr<Mesh>>& m) { return new EMFieldSolver(s, m); }));
    MEMBER(EMFieldSolver, AssembleRhoAndJ, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, AssembleEnergy, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, UpdatePotential, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, RGKStep, void (EMFieldSolver::*)(int, double));
    MEMBER(EMFieldSolver, GetASquared, double (EMFieldSolver::*)(int));
    MEMBER(EMFieldSolver, GetEfield, double (EMFieldSolver::*)(int));
    MEMBER(EMFieldSolver, GetCellAverageASquared, double (EMFieldSolver::*)(int));
    MEMBER(EMFieldSolver, GetMagneticForce, double (EMFieldSolver::*)(int));
    MEMBER(EMFieldSolver, EstimateCFLBound, double (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, EnforceChargeNeutralization, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpCharge, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpEnergy, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpEFieldLongitudinal, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpEFieldTransverse, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpPotential, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpBFieldTransverse, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpAsqField, void (EMFieldSolver::*)());
    MEMBER(EMFieldSolver, DumpTime, void (EMFieldSolver::*)(double));
}
