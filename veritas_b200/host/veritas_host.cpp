// veritas_b200 host layer implementation (see veritas_host.hpp).  Host-side logic only: hierarchy bookkeeping, regridding,
// user callbacks, text output.  Every numerical step of the advance is a call into the C ABI (GPU).
#include "veritas_host.hpp"
#include <cstdlib>

// =====================================================================================================================
// Settings (Settings.cpp:5-195)
// =====================================================================================================================
Settings::Settings(const Input& grid, const Particles& particles, const Output& out) {
    output = out;
    dx = grid.dx; dp = particles.dp;
    maxDepth = (int)grid.Lfinest - 1;
    x_size = grid.nx; p_size = particles.np;
    refinementRatio = grid.r;
    refinementCriteria = grid.refinementCriteria;
    minEfficiency = grid.minEfficiency;
    time = 0.0;
    preLength = grid.preLength; postLength = grid.postLength;
    cfl = grid.cfl; sizeWeight = grid.sizeWeight;
    m = particles.mass; q = particles.charge;
    tempEM = grid.tempEM; temp = particles.misc; pmin = particles.pmin;
    plasma_xl_bound = grid.plasma_xl_bound; plasma_xr_bound = grid.plasma_xr_bound;
    quadratureDepth = 0;

    settingsOverride();      // user hook (veritas.cpp:36-74)

    // fatal configuration errors: message + exit, as the reference does (Settings.cpp:35-50)
    auto fatal = [](const std::string& msg) { std::cerr << "ERROR: " << msg << std::endl; exit(EXIT_FAILURE); };
    if (refinementRatio % 2) fatal("Mesh refinement ratio 'refinementRatio' must be an even value.");
    if (x_size % refinementRatio != 0) fatal("'x_size' must be completely divisible by the Mesh refinement ratio 'refinementRatio'.");
    // the regrid data path fills a one-coarse-cell ring = r fine ghost cells, and a patch has two ghost layers: the reference's own
    // transfer code carries the same restriction (Rectangle.cpp:892-918); refuse instead of writing out of bounds
    if (maxDepth >= 1 && refinementRatio != 2) fatal("a refined mesh (Lfinest > 1) needs 'refinementRatio' = 2: the regrid data transfer is written for that ratio.");
    for (unsigned int s = 0; s < p_size.size(); s++)
        if (p_size[s] % refinementRatio != 0)
            fatal("'p_size' at index (" + std::to_string(s) + ") must be completely divisible by the Mesh refinement ratio 'refinementRatio'.");

    x_size_finest = x_size * std::pow(refinementRatio, maxDepth);
    for (unsigned int s = 0; s < p_size.size(); s++) p_size_finest.push_back(p_size[s] * std::pow(refinementRatio, maxDepth));
    m_inv.resize(m.size());
    for (size_t s = 0; s < m.size(); s++) m_inv[s] = 1 / m[s];
    fMax.resize(q.size(), 0.0);
    DetermineMaximum();

    // run banner (the reference prints a longer one, Settings.cpp:61-125; the viewer only needs the sizes)
    title(" veritas_b200 ", '=');
    std::cout << std::setfill(' ') << "  " << vrt_version() << "\n"
              << "  coarsest grid x_size = " << x_size << ", levels = " << maxDepth + 1 << ", refinement ratio = " << refinementRatio
              << ", finest x_size = " << x_size_finest << "\n";
    for (unsigned int s = 0; s < p_size.size(); s++)
        std::cout << "  species " << s << ": p_size = " << p_size[s] << " (finest " << p_size_finest[s] << "), m = " << m[s] << ", q = " << q[s] << "\n";
    title("", '=');
}

Settings::~Settings() {
    if (gpu_) vrt_destroy(gpu_);
}

void Settings::title(std::string text, char spacer) {
    int space = OUTW - (int)text.length();
    if (space < 2) { std::cout << text << std::endl; return; }
    std::cout << std::string(space / 2, spacer) << text << std::string(space - space / 2, spacer) << std::endl;
}

double Settings::GetDp(int level, int particleType) { return std::pow(refinementRatio, level - maxDepth) * dp.at(particleType); }
double Settings::GetDx(int level) { return std::pow(refinementRatio, level - maxDepth) * dx; }
int Settings::GetXSize(int level) { return x_size * std::pow(refinementRatio, -level + maxDepth); }
int Settings::GetPSize(int level, int particleType) { return p_size[particleType] * std::pow(refinementRatio, -level + maxDepth); }
double Settings::GetMass(int i) { return m.at(i); }
double Settings::GetCharge(int i) { return q.at(i); }
double Settings::GetfMax(int i) { return fMax.at(i); }

void Settings::UpdateTime(int step, double dt) { time = vrt_update_time(time, step, dt); }   // Settings.cpp:166-179

// Settings::DetermineMaximum (Settings.cpp:181-195): running maximum of the initial distribution at p = 0 on the finest x grid;
// the maximum is NOT reset between species (mValue is declared outside the species loop)
void Settings::DetermineMaximum() {
    const double dxf = GetDx(0);
    double mValue = 0.0;
    for (unsigned int s = 0; s < q.size(); s++) {
        for (int i = 0; i < GetXSize(0); i++) {
            const double v = InitialDistribution((i + 0.5) * dxf, 0, s);
            mValue = v > mValue ? v : mValue;
        }
        fMax[s] = mValue;
    }
}

int Settings::PrePad() { return (int)std::max(std::floor(preLength / GetDx(0)), 2.0); }
int Settings::PostPad() { return (int)std::max(std::floor(postLength / GetDx(0)), 2.0); }

void Settings::Check(int rc, const char* what) {
    if (rc == 0) return;
    std::cerr << "ERROR: " << what << " failed (" << rc << "): " << (gpu_ ? vrt_last_error(gpu_) : vrt_global_error()) << std::endl;
    exit(EXIT_FAILURE);
}

vrt_ctx* Settings::Gpu() {
    if (gpu_) return gpu_;
    if (const char* e = getenv("VRT_DEVICE")) device_ = atoi(e);
    Check(vrt_create(&gpu_, device_, (int)q.size()), "vrt_create");
    Check(vrt_set_grid(gpu_, (int)x_size_finest, GetDx(0), PrePad(), PostPad(), (int)refinementRatio, maxDepth), "vrt_set_grid");
    for (unsigned int s = 0; s < q.size(); s++) Check(vrt_set_species(gpu_, (int)s, m[s], q[s], pmin[s], GetDp(0, s)), "vrt_set_species");
    Check(vrt_set_path(gpu_, VRT_PATH_AUTO), "vrt_set_path");
    return gpu_;
}

// =====================================================================================================================
// Rectangle: host descriptor, initial condition, error flags, old -> new data transfer
// =====================================================================================================================
Rectangle::Rectangle(int n_x, int n_p, int x_pos, int p_pos, int depth, Settings& settings, const std::shared_ptr<Rectangle>&, bool up, bool down,
                     bool left, bool right, int particleType)
    : refinementRatio(settings.refinementRatio), depth(depth), particleType(particleType), dp(settings.GetDp(depth, particleType)),
      dx(settings.GetDx(depth)), m_inv(1 / settings.GetMass(particleType)), settings_(&settings), up(up), down(down), left(left), right(right),
      n_x(n_x), n_p(n_p), x_pos(x_pos), p_pos(p_pos), relativeToBottom(std::pow(refinementRatio, depth)) {
    const size_t npad = (size_t)(n_x + 4) * (n_p + 4);
    f.assign(3 * npad, 0.0);
    is_interpolated.assign(npad, false);
    chargeR.assign((size_t)(n_x * relativeToBottom), 0.0);
    currentR.assign((size_t)(n_x * relativeToBottom), 0.0);
    energyR.assign((size_t)(n_p * relativeToBottom), 0.0);
    // error-estimator weights (Rectangle.cpp:70-74): first/second differences in x and p, plus a size term
    const double fm = settings.GetfMax(particleType), r = refinementRatio;
    errorWeights = {0.5 / fm * std::pow(r, -0.5 * depth), 0.5 / fm * std::pow(r, -0.5 * depth), 0.5 / fm * std::pow(r, -2.0 * depth),
                    0.5 / fm * std::pow(r, -2.0 * depth), settings.sizeWeight / fm * std::pow(r, -settings.maxDepth + depth)};
    // conservative cubic through five cell averages (Rectangle.cpp:76-101)
    interpolationMatrix = {0.104166666666667, -0.708333333333334, 0.708333333333334, -0.104166666666667,
                           0.117647058823529, 0.029411764705882,  0.029411764705882, 0.117647058823529,
                           -0.083333333333333, 0.166666666666667, -0.166666666666667, 0.083333333333334};
    interpolatCoefsREF.assign(3 * refinementRatio, 0.0);
    for (int i = 0; i < refinementRatio; i++) {
        const double tl = -0.5 + i / (double)refinementRatio, tr = -0.5 + (i + 1.0) / (double)refinementRatio;
        interpolatCoefsREF[3 * i] = (tl + tr) * 0.5;
        interpolatCoefsREF[3 * i + 1] = (tl * tl + tl * tr + tr * tr) / 3.0 - (1.0 / 12);
        interpolatCoefsREF[3 * i + 2] = (tl * tl * tl + tl * tl * tr + tl * tr * tr + tr * tr * tr) * 0.25;
    }
}
Rectangle::Rectangle() : refinementRatio(0), depth(0), particleType(0), dp(0), dx(0), m_inv(0), settings_(nullptr), up(false), down(false),
                         left(false), right(false), n_x(0), n_p(0), x_pos(0), p_pos(0), relativeToBottom(0) {}
Rectangle::~Rectangle() {}
BoundaryCondition::~BoundaryCondition() {}

double Rectangle::GetValueFromSameLevel(int i, int j, int val) { return f[Index3(i - x_pos, j - p_pos, val)]; }
double BoundaryCondition::GetValueFromSameLevel(int, int, int) { return 0.0; }   // BoundaryCondition.cpp:6-8

vrt_patch_desc Rectangle::Descriptor() const {
    vrt_patch_desc d;
    d.depth = depth; d.x_pos = x_pos; d.p_pos = p_pos; d.n_x = n_x; d.n_p = n_p;
    d.up = up; d.down = down; d.left = left; d.right = right;
    return d;
}

// Rectangle::InitializeDistribution (Rectangle.cpp:616-665, MKLINIT == 0 branch): sub-cell midpoint quadrature of the user's
// initial distribution, nvals = r^(depth + quadratureDepth) points per direction; states 0 and 1
void Rectangle::InitializeDistribution() {
    const int nvals = std::pow(refinementRatio, depth + settings_->quadratureDepth);
#pragma omp parallel for schedule(guided)
    for (int i = 0; i < n_x; i++)
        for (int j = 0; j < n_p; j++) {
            double acc = 0.0;
            for (int k = 0; k < nvals; k++)
                for (int l = 0; l < nvals; l++) {
                    const double xq = ((0.5 + k) / nvals + i + x_pos) * dx;
                    const double pq = Momentum((0.5 + l) / nvals + j);
                    acc += settings_->InitialDistribution(xq, pq, particleType);
                }
            const double v = acc * (1 / std::pow(nvals, 2.0));
            f[Index3(i, j, 0)] = v;
            f[Index3(i, j, 1)] = v;
        }
}

// Rectangle::ErrorEstimate (Rectangle.hpp:128-130) on state 1
double Rectangle::ErrorEstimate(int i, int j) {
    const double c = f[Index3(i, j, 1)], xp = f[Index3(i + 1, j, 1)], xm = f[Index3(i - 1, j, 1)], pp = f[Index3(i, j + 1, 1)], pm = f[Index3(i, j - 1, 1)];
    return errorWeights[0] * std::fabs(xp - xm) + errorWeights[1] * std::fabs(pp - pm) + errorWeights[2] * std::fabs(xp - 2 * c + xm) +
           errorWeights[3] * std::fabs(pp - 2 * c + pm) + errorWeights[4] * std::fabs(c);
}

// Rectangle::getError (Rectangle.cpp:866-890): cells of this patch (global indices of its level) whose error estimate exceeds the
// criterion or which the user forces; a margin of 3 * sum_{i<=maxDepth-depth} r^i cells at the domain edge is never flagged
void Rectangle::getError(std::vector<coords>& flaggedCells, int particleType) {
    int margin = 0;
    for (int i = 0; i <= settings_->maxDepth - depth; i++) margin += std::pow(settings_->refinementRatio, i);
    margin *= 3;
    const int lo_x = std::max(x_pos, margin), hi_x = std::min(x_pos + n_x, settings_->GetXSize(depth) - margin);
    const int lo_p = std::max(p_pos, margin), hi_p = std::min(p_pos + n_p, settings_->GetPSize(depth, particleType) - margin);
    // the estimate is evaluated where the data live: on the device (one byte per cell comes back) or on the host mirror
    std::vector<unsigned char> dev;
    if (device_flags_) {
        dev.resize((size_t)n_x * n_p);
        settings_->Check(vrt_error_flags(settings_->Gpu(), particleType, patch_id, errorWeights.data(), settings_->refinementCriteria, dev.data()),
                         "vrt_error_flags");
    }
    for (int i = lo_x; i < hi_x; i++)
        for (int j = lo_p; j < hi_p; j++) {
            const bool over = device_flags_ ? dev[(size_t)(i - x_pos) * n_p + (j - p_pos)] != 0
                                            : ErrorEstimate(i - x_pos, j - p_pos) > settings_->refinementCriteria;
            if (over || settings_->RefinementOverride((i + 0.5) * dx, Momentum(j - p_pos), depth, particleType))
                flaggedCells.push_back(std::make_pair(i, j));
        }
}

// Rectangle::GetInterpolantsREF (Rectangle.cpp:121-137)
std::vector<double> Rectangle::GetInterpolantsREF(double f1, double f2, double f3, double f4, double f5) {
    std::vector<double> out(refinementRatio, 0.0);
    f5 -= f3; f4 -= f3; f2 -= f3; f1 -= f3;
    const std::vector<double>& M = interpolationMatrix;
    const double a1 = M[0] * f1 + M[1] * f2 + M[2] * f4 + M[3] * f5;
    const double a2 = M[4] * f1 + M[5] * f2 + M[6] * f4 + M[7] * f5;
    const double a3 = M[8] * f1 + M[9] * f2 + M[10] * f4 + M[11] * f5;
    for (int i = 0; i < refinementRatio; i++)
        out[i] = interpolatCoefsREF[3 * i] * a1 + interpolatCoefsREF[3 * i + 1] * a2 + interpolatCoefsREF[3 * i + 2] * a3 + f3;
    return out;
}

// Rectangle::GetWenoValueFromCoarseLevel (Rectangle.cpp:343-415), host version used by the regrid data transfer (d = -1:
// all r*r sub-cells, x-major); the ghost-fill variants (d = 0..3) run on the device (k_ghost_coarse)
std::vector<double> Rectangle::GetWenoValueFromCoarseLevel(int i, int j, int d, int val) {
    const int r = refinementRatio, ic = i / r - x_pos, jc = j / r - p_pos;
    std::vector<std::vector<double>> rows;
    for (int k = -2; k < 3; k++)
        rows.push_back(GetInterpolantsREF(f[Index3(ic - 2, jc + k, val)], f[Index3(ic - 1, jc + k, val)], f[Index3(ic, jc + k, val)],
                                          f[Index3(ic + 1, jc + k, val)], f[Index3(ic + 2, jc + k, val)]));
    std::vector<double> sub(r * r);
    double sum = 0.0;
    for (int k = 0; k < r; k++) {
        const std::vector<double> part = GetInterpolantsREF(rows[0][k], rows[1][k], rows[2][k], rows[3][k], rows[4][k]);
        for (int l = 0; l < r; l++) { sub[k * r + l] = part[l]; sum += part[l]; }
    }
    const double correction = f[Index3(ic, jc, val)] - 1.0 / std::pow(r, 2) * sum;
    for (double& v : sub) v += correction;
    if (d < 0 || d > 3) return sub;
    std::vector<double> two(2 * r);
    for (int k = 0; k < r; k++) {
        if (d == 0) { two[2 * k] = sub[r * k + r - 1]; two[2 * k + 1] = sub[r * k + r - 2]; }
        else if (d == 1) { two[2 * k] = sub[k]; two[2 * k + 1] = sub[r + k]; }
        else if (d == 2) { two[2 * k] = sub[r * k]; two[2 * k + 1] = sub[r * k + 1]; }
        else { two[2 * k] = sub[r * (r - 1) + k]; two[2 * k + 1] = sub[r * (r - 2) + k]; }
    }
    return two;
}

// old coarse patch (this) -> new finer patch: interpolate every coarse cell under the target plus one ring (Rectangle.cpp:892-918)
void Rectangle::GetDataFromCoarseLevelRectangle(const std::shared_ptr<Rectangle>& target) {
    const int r = refinementRatio;
    const int tx = target->x_pos / r, tp = target->p_pos / r, tnx = target->n_x / r, tnp = target->n_p / r;
    const int lo_x = std::max(x_pos, tx - 1), hi_x = std::min(x_pos + n_x, tx + tnx + 1);
    const int lo_p = std::max(p_pos, tp - 1), hi_p = std::min(p_pos + n_p, tp + tnp + 1);
    for (int i = lo_x; i < hi_x; i++)
        for (int j = lo_p; j < hi_p; j++) {
            const std::vector<double> sub = GetWenoValueFromCoarseLevel(r * i + 1, r * j + 1, -1);
            for (int k = 0; k < r; k++)
                for (int l = 0; l < r; l++) {
                    const int ti = (i - tx) * r + k, tj = (j - tp) * r + l;
                    target->f[target->Index3(ti, tj, 0)] = sub[l + r * k];
                    target->f[target->Index3(ti, tj, 1)] = target->f[target->Index3(ti, tj, 0)];
                    target->is_interpolated[target->IndexNS(ti, tj)] = true;
                }
        }
}
// old same-level patch (this) -> new patch: copy the overlap, the target's two ghost layers included (Rectangle.cpp:920-941)
void Rectangle::GetDataFromSameLevelRectangle(const std::shared_ptr<Rectangle>& target) {
    const int tx = target->x_pos, tp = target->p_pos, tnx = target->n_x, tnp = target->n_p;
    const int lo_x = std::max(x_pos, tx - 2), hi_x = std::min(x_pos + n_x, tx + tnx + 2);
    const int lo_p = std::max(p_pos, tp - 2), hi_p = std::min(p_pos + n_p, tp + tnp + 2);
    for (int i = lo_x; i < hi_x; i++)
        for (int j = lo_p; j < hi_p; j++) {
            target->f[target->Index3(i - tx, j - tp, 0)] = GetValueFromSameLevel(i, j);
            target->f[target->Index3(i - tx, j - tp, 1)] = target->f[target->Index3(i - tx, j - tp, 0)];
            target->is_interpolated[target->IndexNS(i - tx, j - tp)] = true;
        }
}
// new coarse patch (this) -> new finer patch, only where no old data reached the target (Rectangle.cpp:1100-1128)
void Rectangle::GetDataFromCoarseNewLevelRectangle(const std::shared_ptr<Rectangle>& target) {
    const int r = refinementRatio;
    const int tx = target->x_pos / r, tp = target->p_pos / r, tnx = target->n_x / r, tnp = target->n_p / r;
    const int lo_x = std::max(x_pos, tx - 1), hi_x = std::min(x_pos + n_x, tx + tnx + 1);
    const int lo_p = std::max(p_pos, tp - 1), hi_p = std::min(p_pos + n_p, tp + tnp + 1);
    for (int i = lo_x; i < hi_x; i++)
        for (int j = lo_p; j < hi_p; j++) {
            if (target->is_interpolated[target->IndexNS(r * (i - tx) + 1, r * (j - tp) + 1)]) continue;
            const std::vector<double> sub = GetWenoValueFromCoarseLevel(r * i + 1, r * j + 1, -1);
            for (int k = 0; k < r; k++)
                for (int l = 0; l < r; l++) {
                    const int ti = (i - tx) * r + k, tj = (j - tp) * r + l;
                    target->f[target->Index3(ti, tj, 0)] = sub[l + r * k];
                    target->f[target->Index3(ti, tj, 1)] = target->f[target->Index3(ti, tj, 0)];
                }
        }
}

// =====================================================================================================================
// Level
// =====================================================================================================================
Level::Level(int particleType, int depth, Settings& settings)
    : x_size(settings.GetXSize(settings.maxDepth - depth)), p_size(settings.GetPSize(settings.maxDepth - depth, particleType)), depth(depth),
      particleType(particleType), settings(settings) {}

// Level::FCTTimeStep (Level.cpp:12-17): one launch set for all rectangles of the level
void Level::FCTTimeStep(double timestep, int step, int subStep) {
    if (rectangles.empty()) return;
    settings.Check(vrt_vlasov_substep(settings.Gpu(), particleType, depth, timestep, step, subStep), "vrt_vlasov_substep");
}
// Level::PushData (Level.cpp:88-126)
void Level::PushData(int updateType, int val) {
    if (rectangles.empty()) return;
    settings.Check(vrt_level_push(settings.Gpu(), particleType, depth, updateType, val), "vrt_level_push");
}
// Level::CollectRhoAndJ (Level.cpp:42-62): Rectangle::CalculateRhoAndJ on the device (vrt_patch_moments fills chargeR, currentR),
// level sum on the host in the reference's order.  SolverManager::Advance does not come through here: EMFieldSolver::AssembleRhoAndJ
// is one device pass over all levels and species (vrt_moments).
void Level::CollectRhoAndJ() {
    chargeL.assign(settings.x_size_finest, 0.0);
    currentL.assign(settings.x_size_finest, 0.0);
    for (auto& r : rectangles) {
        settings.Check(vrt_patch_moments(settings.Gpu(), particleType, r->patch_id, r->chargeR.data(), r->currentR.data()), "vrt_patch_moments");
        const int shift = r->x_pos, rtb = (int)r->relativeToBottom;
        for (size_t j = 0; j < r->chargeR.size(); j++) { chargeL[shift * rtb + j] += r->chargeR[j]; currentL[shift * rtb + j] += r->currentR[j]; }
    }
}
void Level::InterpolateRhoAndJToFinestMesh(std::vector<double>& charge, std::vector<double>& J) {     // Level.cpp:19-29
    CollectRhoAndJ();
    for (unsigned int i = 0; i < settings.x_size_finest; i++) { charge[i] += chargeL[i]; J[i] += currentL[i]; }
}
// Level::CollectEnergy (Level.cpp:64-78): Rectangle::CalculateEnergy on the device (vrt_patch_energy), level sum on the host
void Level::CollectEnergy() {
    energyL.assign(settings.p_size_finest[particleType], 0.0);
    for (auto& r : rectangles) {
        settings.Check(vrt_patch_energy(settings.Gpu(), particleType, r->patch_id, r->energyR.data()), "vrt_patch_energy");
        const int shift = r->p_pos, rtb = r->relativeToBottom;
        for (size_t i = 0; i < r->energyR.size(); i++) energyL[shift * rtb + i] += r->energyR[i];
    }
}
void Level::InterpolateEnergyToFinestMesh(std::vector<double>& energy) {     // Level.cpp:31-40
    CollectEnergy();
    const unsigned int n = settings.p_size_finest[particleType];
    for (unsigned int i = 0; i < n; i++) energy[i] += energyL[i];
}
// this = new level, `level` = the old one (Level.cpp:128-150): every old rectangle feeds every new one
void Level::GetDataFromSameLevel(const std::unique_ptr<Level>& level) {
    for (auto& src : level->rectangles) for (auto& dst : rectangles) src->GetDataFromSameLevelRectangle(dst);
}
void Level::GetDataFromCoarserLevel(const std::unique_ptr<Level>& level) {
    for (auto& src : level->rectangles) for (auto& dst : rectangles) src->GetDataFromCoarseLevelRectangle(dst);
}
void Level::GetDataFromCoarseNewLevel(const std::unique_ptr<Level>& level) {
    for (auto& src : level->rectangles) for (auto& dst : rectangles) src->GetDataFromCoarseNewLevelRectangle(dst);
}

// =====================================================================================================================
// Mesh: hierarchy, regridding, device hand-over
// =====================================================================================================================
// Mesh::Mesh (Mesh.cpp:8-50).  With MKLINIT == 0 the analytic initial flags never reach the caller (quirk Q6), so the
// constructor starts from the base level alone and grows the hierarchy from data by maxDepth calls of updateHierarchy(true).
Mesh::Mesh(int particleType, Settings& settings) : settings(settings), EMSolver(nullptr), particleType(particleType) {
    bc = std::make_shared<BoundaryCondition>();
    std::vector<coords> flagged;
    for (int l = 0; l <= settings.maxDepth; l++) {
        if (l == 0) {
            level base;
            base.push_back(std::make_pair(std::make_pair(0, 0), std::make_pair(settings.GetXSize(settings.maxDepth), settings.GetPSize(settings.maxDepth, particleType))));
            hierarchy.push_back(base);
            getError(l, true, flagged);
        } else {
            if (!flagged.empty()) {
                rect extrema;
                getExtrema(extrema, flagged);
                level found = splitRectangle(extrema, flagged, settings.minEfficiency);
                interpRectanglesUp(found, l - 1);
                hierarchy.push_back(found);
            }
            flagged.clear();
            getError(l, true, flagged);
        }
    }
    promoteHierarchyToMesh(true);
    for (int i = 0; i < settings.maxDepth; i++) {
        PushData();
        updateHierarchy(true);
    }
}

void Mesh::SetFieldSolver(const std::shared_ptr<EMFieldSolver>& solver) { EMSolver = solver; }

// Mesh::Advance (Mesh.cpp:64-89): sub-steps 0..2 (3 at the last stage) with the ghost and limiter syncs between them
void Mesh::Advance(double timeStep, int step) {
    settings.Check(vrt_vlasov_stage(settings.Gpu(), particleType, timeStep, step), "vrt_vlasov_stage");
    device_current_ = true;
}
void Mesh::PushData(int val) {
    settings.Check(vrt_push_data(settings.Gpu(), particleType, val), "vrt_push_data");
    device_current_ = true;
}
void Mesh::PushBoundaryC() { settings.Check(vrt_push_boundary_c(settings.Gpu(), particleType), "vrt_push_boundary_c"); }

void Mesh::InterpolateRhoAndJToFinestMesh(std::vector<double>& charge, std::vector<double>& J) {
    settings.Check(vrt_moments_species(settings.Gpu(), particleType, charge.data(), J.data()), "vrt_moments_species");
}

void Mesh::InterpolateEnergyToFinestMesh(std::vector<double>& energy) {     // Mesh.cpp:58-62
    for (auto& lvl : levels) lvl->InterpolateEnergyToFinestMesh(energy);
}

void Mesh::SyncHost() {
    if (!device_current_) return;
    vrt_ctx* g = settings.Gpu();
    std::vector<double> plane;
    for (auto& lvl : levels)
        for (auto& r : lvl->rectangles) {
            const size_t npad = (size_t)(r->n_x + 4) * (r->n_p + 4);
            plane.resize(npad);
            for (int state = 0; state < 2; state++) {
                settings.Check(vrt_patch_download_f(g, particleType, r->patch_id, state, plane.data()), "vrt_patch_download_f");
                for (size_t c = 0; c < npad; c++) r->f[3 * c + state] = plane[c];
            }
        }
    device_current_ = false;
}

// ---- error flags ---------------------------------------------------------------------------------------------------------
// Mesh::getError (Mesh.cpp:212-295).  init: the analytic pre-flagging; with MKLINIT == 0 its result is discarded (quirk Q6).
void Mesh::getError(const int& lvl, bool init, std::vector<coords>& flaggedCells) {
    if (init) return;
    for (auto& r : levels.at(settings.maxDepth - lvl)->rectangles) { r->device_flags_ = device_current_ && device_regrid_; r->getError(flaggedCells, particleType); }
}

void Mesh::getExtrema(rect& extrema, const std::vector<coords>& flaggedCells) {      // bounding box of the flags (Mesh.cpp:625-642)
    int x0 = flaggedCells.front().first, x1 = x0, p0 = flaggedCells.front().second, p1 = p0;
    for (const coords& c : flaggedCells) {
        x0 = std::min(x0, c.first); x1 = std::max(x1, c.first);
        p0 = std::min(p0, c.second); p1 = std::max(p1, c.second);
    }
    extrema = std::make_pair(std::make_pair(x0, p0), std::make_pair(x1, p1));
}

int Mesh::countCells(const rect& span) { return (span.second.first - span.first.first + 1) * (span.second.second - span.first.second + 1); }

// signatures = number of flags per column / per row of the box, and their discrete Laplacians (Mesh.cpp:386-446)
std::tuple<std::vector<int>, std::vector<int>, std::vector<int>, std::vector<int>> Mesh::computeSignatures(rect& box, std::vector<coords>& flagged) {
    const int lx = box.second.first - box.first.first, lp = box.second.second - box.first.second;
    std::vector<int> sigX(lx + 1, 0), sigP(lp + 1, 0), lapX, lapP;
    for (const coords& c : flagged) { sigX.at(c.first - box.first.first)++; sigP.at(c.second - box.first.second)++; }
    if (lx >= 2) { lapX.resize(lx - 1); for (int h = 1; h < lx; h++) lapX[h - 1] = sigX[h - 1] + sigX[h + 1] - 2 * sigX[h]; }
    if (lp >= 2) { lapP.resize(lp - 1); for (int h = 1; h < lp; h++) lapP[h - 1] = sigP[h - 1] + sigP[h + 1] - 2 * sigP[h]; }
    return std::make_tuple(sigX, lapX, sigP, lapP);
}

// first zero of a signature and the extent of that gap (Mesh.cpp:594-623): returns (found, last index kept by piece one,
// first index of piece two), both relative to the signature
std::tuple<bool, int, int> Mesh::hasHole(std::vector<int>& sig) {
    const int n = (int)sig.size();
    int z = 0;
    while (z < n && sig[z] != 0) z++;
    if (z == n) return std::make_tuple(false, 0, 0);
    const int start = (z == 0) ? 0 : z - 1;
    int e = z;
    while (e < n && sig[e] == 0) e++;
    const int finish = (e == n) ? n - 1 : e;
    return std::make_tuple(true, start, finish);
}

namespace {
inline int sign_of(int x) { return (x > 0) - (x < 0); }
// strongest sign change of a Laplacian signature (Mesh.cpp:455-515): returns (index, strength, must_bisect)
struct Inflection { int index = 0, strength = 0; bool bisect = false; };
Inflection strongest_inflection(const std::vector<int>& lap, int fallback_range) {
    Inflection out;
    const int n = (int)lap.size();
    if (n <= 2) return out;
    std::vector<int> at, size;
    for (int i = 0; i + 1 < n; i++)
        if (lap[i] != 0 && lap[i + 1] != 0 && sign_of(lap[i]) != sign_of(lap[i + 1])) { at.push_back(i); size.push_back(std::abs(lap[i + 1] - lap[i])); }
    if (at.empty()) { out.bisect = true; return out; }
    if (at.size() == 1) { out.index = at[0]; out.strength = size[0]; return out; }
    const auto mx = std::max_element(size.begin(), size.end());
    if (std::count(size.begin(), size.end(), *mx) > 1) {
        // several equally strong candidates: the one nearest the centre — among ALL sign changes, not only the strongest
        // ones, and the first on ties (reference behaviour)
        const int centre = n / 2;
        int range = fallback_range, chosen = 0;
        for (size_t i = 0; i < at.size(); i++)
            if (range > std::abs(at[i] - centre)) { range = std::abs(at[i] - centre); chosen = (int)i; }
        out.index = at[chosen]; out.strength = size[chosen];
    } else {
        out.index = at[mx - size.begin()]; out.strength = size[mx - size.begin()];
    }
    return out;
}
}  // namespace

// where to cut a box without holes (Mesh.cpp:448-592): (0, x) cut after column x, (1, p) cut after row p, (2, -) bisect the
// long edge, (3, -) leave it
coords Mesh::identifyInflection(rect& box, std::tuple<std::vector<int>, std::vector<int>, std::vector<int>, std::vector<int>>& signatures) {
    const std::vector<int>& lapX = std::get<1>(signatures);
    const std::vector<int>& lapP = std::get<3>(signatures);
    const Inflection ix = strongest_inflection(lapX, std::numeric_limits<int>::max());
    const Inflection ip = strongest_inflection(lapP, 32767);
    if (ix.bisect && ip.bisect) return std::make_pair(2, 0);
    if ((lapX.size() <= 2 && lapP.size() <= 2) || (ix.strength == 0 && ip.strength == 0)) return std::make_pair(3, 0);
    if (ix.strength > ip.strength) return std::make_pair(0, ix.index + box.first.first + 1);
    return std::make_pair(1, ip.index + box.first.second + 1);
}

// Berger-Rigoutsos style clustering (Mesh.cpp:644-792): accept the box if its fill ratio exceeds minEfficiency; otherwise cut at
// a hole of a signature, else at the strongest inflection, else bisect; recurse into both pieces
level Mesh::splitRectangle(rect& box, std::vector<coords>& flagged, const double& minEfficiency) {
    level result;
    if (flagged.size() / double(countCells(box)) > minEfficiency) { result.push_back(box); return result; }
    auto sig = computeSignatures(box, flagged);
    std::tuple<bool, int, int> hx = hasHole(std::get<0>(sig)), hp = hasHole(std::get<2>(sig));
    std::get<1>(hx) += box.first.first; std::get<2>(hx) += box.first.first;
    std::get<1>(hp) += box.first.second; std::get<2>(hp) += box.first.second;
    rect one, two;
    auto cut_x = [&](int last_of_one, int first_of_two) {
        one = std::make_pair(box.first, coords(last_of_one, box.second.second));
        two = std::make_pair(coords(first_of_two, box.first.second), box.second);
    };
    auto cut_p = [&](int last_of_one, int first_of_two) {
        one = std::make_pair(box.first, coords(box.second.first, last_of_one));
        two = std::make_pair(coords(box.first.first, first_of_two), box.second);
    };
    if (std::get<0>(hx) && std::get<0>(hp)) {
        if (std::abs(std::get<2>(hx) - std::get<1>(hx)) > std::abs(std::get<2>(hp) - std::get<1>(hp))) cut_x(std::get<1>(hx), std::get<2>(hx));
        else cut_p(std::get<1>(hp), std::get<2>(hp));
    } else if (std::get<0>(hx)) {
        cut_x(std::get<1>(hx), std::get<2>(hx));
    } else if (std::get<0>(hp)) {
        cut_p(std::get<1>(hp), std::get<2>(hp));
    } else {
        const coords where = identifyInflection(box, sig);
        if (where.first == 0) cut_x(where.second, where.second + 1);
        else if (where.first == 1) cut_p(where.second, where.second + 1);
        else if (where.first == 2) {
            const double lx = box.second.first - box.first.first, lp = box.second.second - box.first.second;
            if (lx > lp) { const int c = (int)std::floor((lx / 2) + box.first.first); cut_x(c, c + 1); }
            else { const int c = (int)std::floor((lp / 2) + box.first.second); cut_p(c, c + 1); }
        } else { result.push_back(box); return result; }
    }
    std::vector<coords> in_one, in_two;
    for (const coords& c : flagged) {
        const bool inside = c.first >= one.first.first && c.first <= one.second.first && c.second >= one.first.second && c.second <= one.second.second;
        (inside ? in_one : in_two).push_back(c);
    }
    level rest;
    if (!in_one.empty()) result = splitRectangle(one, in_one, minEfficiency);
    if (!in_two.empty()) rest = splitRectangle(two, in_two, minEfficiency);
    result.insert(result.end(), rest.begin(), rest.end());
    return result;
}

// boxes found on hierarchy level lvl become patches of level lvl+1: inclusive upper corner -> exclusive, then scale (Mesh.cpp:298-313)
void Mesh::scaleRectanglesUp(level& identified, int nmax, int pmax, int nmax1, int pmax1) {
    for (rect& r : identified) {
        r.second.first += 1; r.second.second += 1;
        r.first.first *= (nmax1 / double(nmax)); r.first.second *= (pmax1 / double(pmax));
        r.second.first *= (nmax1 / double(nmax)); r.second.second *= (pmax1 / double(pmax));
    }
}
void Mesh::interpRectanglesUp(level& identified, const int& lvl) {
    scaleRectanglesUp(identified, settings.GetXSize(settings.maxDepth - lvl), settings.GetPSize(settings.maxDepth - lvl, particleType),
                      settings.GetXSize(settings.maxDepth - (lvl + 1)), settings.GetPSize(settings.maxDepth - (lvl + 1), particleType));
}

// cells of level lvl under a patch of level lvl+2 widened by 2r cells (one more on the upper x side), so that the regridded
// level lvl+1 keeps containing it (Mesh.cpp:315-339); the index scaling truncates towards zero as the reference's int conversion does
void Mesh::footprintBelow(const rect& r, int nmax, int pmax, int nmax2, int pmax2, int ratio, std::vector<coords>& foundCells) {
    const int w = 2 * ratio;
    for (int i = r.first.first - w; i <= r.second.first + w + 1; i++)
        for (int j = r.first.second - w; j <= r.second.second + w; j++) {
            const int ic = i * (nmax / double(nmax2)), jc = j * (pmax / double(pmax2));
            if (ic > -1 && ic < nmax && jc > -1 && jc < pmax) foundCells.push_back(std::make_pair(ic, jc));
        }
    std::sort(foundCells.begin(), foundCells.end());
    foundCells.erase(std::unique(foundCells.begin(), foundCells.end()), foundCells.end());
}
void Mesh::mergeDownFlaggedData(const int& lvl, const rect& r, std::vector<coords>& foundCells) {
    footprintBelow(r, settings.GetXSize(settings.maxDepth - lvl), settings.GetPSize(settings.maxDepth - lvl, particleType),
                   settings.GetXSize(settings.maxDepth - (lvl + 2)), settings.GetPSize(settings.maxDepth - (lvl + 2), particleType),
                   (int)settings.refinementRatio, foundCells);
}

// Mesh::updateHierarchy (Mesh.cpp:132-210): from the finest hierarchy level down, flag -> cluster -> replace the next finer
// level; flags of level l are united with the footprint of the level-(l+2) patches so that nesting survives
void Mesh::updateHierarchy(bool init) {
    if (hierarchy.empty()) { std::cerr << "Exception Occured: hierarchy is empty, cannot update it." << std::endl; exit(EXIT_FAILURE); }
    // VRT_HOST_REGRID=1 keeps the reference's host-side regrid data path (download, Rectangle::getError and
    // InterMeshDataTransfer on the mirrors, upload); default: error flags and old -> new transfer run on the device
    device_regrid_ = !(getenv("VRT_HOST_REGRID") && atoi(getenv("VRT_HOST_REGRID")));
    if (!device_regrid_) SyncHost();
    std::vector<coords> flagged, own, below, tmp;
    const int top = (int)hierarchy.size() - 1;
    for (int l = top; l > -1; l--) {
        if (l < (int)hierarchy.size() - 2) {
            own.clear(); below.clear();
            if (hierarchy.size() > 2)
                for (rect& r : hierarchy.at(l + 2)) {
                    tmp.clear();
                    mergeDownFlaggedData(l, r, tmp);
                    below.insert(below.end(), tmp.begin(), tmp.end());
                }
            getError(l, false, own);
            std::sort(own.begin(), own.end());
            flagged.reserve(own.size() + below.size());
            std::set_union(own.begin(), own.end(), below.begin(), below.end(), std::back_inserter(flagged));
        } else {
            getError(l, false, flagged);
        }
        if (!flagged.empty()) {
            rect extrema;
            getExtrema(extrema, flagged);
            level found = splitRectangle(extrema, flagged, settings.minEfficiency);
            if (l == top) {
                if (l < settings.maxDepth) { interpRectanglesUp(found, l); hierarchy.push_back(found); }
            } else {
                interpRectanglesUp(found, l);
                hierarchy.at(l + 1) = found;
            }
        } else if (l < top) {
            hierarchy.erase(hierarchy.begin() + l + 1);
        }
        flagged.clear();
    }
    promoteHierarchyToMesh(init);
}

void Mesh::InterMeshDataTransfer(const std::vector<std::unique_ptr<Level>>& old_levels) {     // Mesh.cpp:116-130
    for (unsigned int i = 0; i + 1 < old_levels.size(); i++) {
        levels[i]->GetDataFromCoarserLevel(old_levels[i + 1]);
        levels[i]->GetDataFromSameLevel(old_levels[i]);
    }
    if (old_levels.size() > 1) levels[old_levels.size() - 1]->GetDataFromSameLevel(old_levels[old_levels.size() - 1]);
    for (int i = (int)levels.size() - 1; i > 0; i--) levels[i - 1]->GetDataFromCoarseNewLevel(levels[i]);
}

// Mesh::promoteHierarchyToMesh (Mesh.cpp:794-875): Level / Rectangle objects for the hierarchy, data (initial condition or
// transfer from the old mesh), then the hand-over to the device: descriptors -> vrt_set_hierarchy (connectivity is derived
// there), f -> vrt_patch_upload_f, PushData, FCTTimeStep(0,-1,3) -> vrt_commit_state
void Mesh::promoteHierarchyToMesh(bool init) {
    const int N = (int)hierarchy.size() - 1, empty = settings.maxDepth - N;
    vrt_ctx* g = settings.Gpu();
    const bool on_device = !init && device_regrid_ && device_current_ && settings.maxDepth >= 1 &&
                           vrt_get_path(g, particleType) == VRT_PATH_SPLIT;
    if (!init && !on_device) SyncHost();          // the host transfer below reads the old rectangles' mirrors
    std::vector<std::unique_ptr<Level>> old_levels;
    if (!init) for (auto& lvl : levels) old_levels.push_back(std::move(lvl));
    levels.clear();
    for (int i = 0; i < empty; i++) levels.push_back(std::make_unique<Level>(particleType, i, settings));
    for (int lvl = N; lvl >= 0; lvl--) {
        // quirk Q5: the domain extent is taken at depth N - lvl, not N - lvl + empty
        const int xmax = settings.GetXSize(N - lvl), pmax = settings.GetPSize(N - lvl, particleType);
        const int depth = N - lvl + empty;
        levels.push_back(std::make_unique<Level>(particleType, depth, settings));
        for (rect& r : hierarchy.at(lvl))
            levels.at(depth)->rectangles.push_back(std::make_shared<Rectangle>(
                r.second.first - r.first.first, r.second.second - r.first.second, r.first.first, r.first.second, depth, settings, bc,
                r.second.second == pmax, r.first.second == 0, r.first.first == 0, r.second.first == xmax, particleType));
    }
    std::vector<vrt_patch_desc> desc;
    int id = 0;
    for (auto& lvl : levels) for (auto& r : lvl->rectangles) { r->patch_id = id++; desc.push_back(r->Descriptor()); }
    if (on_device) {
        // InterMeshDataTransfer between the old and the new device patch tables; Rectangle::f of the new patches is a stale
        // mirror until the next SyncHost()
        settings.Check(vrt_regrid(g, particleType, (int)desc.size(), desc.data()), "vrt_regrid");
    } else {
        if (init) for (auto& lvl : levels) for (auto& r : lvl->rectangles) r->InitializeDistribution();
        if (!init) InterMeshDataTransfer(old_levels);
        // AMR hierarchies run on the split path; a single full-domain patch takes the fused streaming path
        settings.Check(vrt_set_hierarchy(g, particleType, (int)desc.size(), desc.data()), "vrt_set_hierarchy");
        std::vector<double> plane;
        for (auto& lvl : levels)
            for (auto& r : lvl->rectangles) {
                const size_t npad = (size_t)(r->n_x + 4) * (r->n_p + 4);
                plane.resize(npad);
                for (int state = 0; state < 2; state++) {
                    for (size_t c = 0; c < npad; c++) plane[c] = r->f[3 * c + state];
                    settings.Check(vrt_patch_upload_f(g, particleType, r->patch_id, state, plane.data()), "vrt_patch_upload_f");
                }
            }
    }
    PushData();
    settings.Check(vrt_commit_state(g, particleType), "vrt_commit_state");
    device_current_ = true;
}

// After vrt_checkpoint_read the device holds the checkpoint's hierarchy: rebuild the Level / Rectangle objects (same order as
// promoteHierarchyToMesh numbers them: levels by depth, rectangles in stored order) and the level boxes the next regrid starts
// from.  Rectangle::f of the new objects is a stale mirror until the next SyncHost().
void Mesh::AdoptDeviceHierarchy() {
    vrt_ctx* g = settings.Gpu();
    const int n = vrt_get_hierarchy(g, particleType, 0, nullptr);
    if (n < 1) settings.Check(n < 0 ? n : -1, "vrt_get_hierarchy");
    std::vector<vrt_patch_desc> d(n);
    vrt_get_hierarchy(g, particleType, n, d.data());
    levels.clear();
    for (int depth = 0; depth <= settings.maxDepth; depth++) levels.push_back(std::make_unique<Level>(particleType, depth, settings));
    int id = 0, min_depth = settings.maxDepth;
    for (const vrt_patch_desc& q : d) {
        auto r = std::make_shared<Rectangle>(q.n_x, q.n_p, q.x_pos, q.p_pos, q.depth, settings, bc, q.up != 0, q.down != 0, q.left != 0, q.right != 0, particleType);
        r->patch_id = id++;
        levels.at(q.depth)->rectangles.push_back(r);
        min_depth = std::min(min_depth, q.depth);
    }
    hierarchy.assign(settings.maxDepth - min_depth + 1, level());
    for (const vrt_patch_desc& q : d)
        hierarchy.at(settings.maxDepth - q.depth).push_back(std::make_pair(std::make_pair(q.x_pos, q.p_pos), std::make_pair(q.x_pos + q.n_x, q.p_pos + q.n_p)));
    device_current_ = true;
}

// Mesh::outputRectangleData (Mesh.cpp:877-902): one text file per level, "r x_pos p_pos n_x n_p" then "i j f" per cell (state 0)
void Mesh::outputRectangleData(double tidx) {
    SyncHost();
    for (int l = 0; l <= settings.maxDepth; l++) {
        if (levels.at(l)->rectangles.empty()) continue;
        std::stringstream name;
        name << "output/rectangleData/rectangleData_p" << particleType << "_l" << l << "_t" << std::scientific << tidx << ".txt";
        std::ofstream out(name.str());
        out << std::setprecision(4);
        for (auto& r : levels[l]->rectangles) {
            out << "r " << r->x_pos << " " << r->p_pos << " " << r->n_x << " " << r->n_p << "\n";
            for (int i = 0; i < r->n_x; i++)
                for (int j = 0; j < r->n_p; j++)
                    out << r->x_pos + i << " " << r->p_pos + j << " " << std::scientific << " " << r->f[r->Index3(i, j, 0)] << "\n";
        }
    }
}

// =====================================================================================================================
// EMFieldSolver: host mirrors of the device-resident 1-D solver
// =====================================================================================================================
EMFieldSolver::EMFieldSolver(Settings& settings, const std::vector<std::shared_ptr<Mesh>>& meshes)
    : settings(settings), n_prepad(settings.PrePad()), n_postpad(settings.PostPad()), x_size(settings.x_size_finest), meshes(meshes),
      chargeStream(nullptr), ELongStream(nullptr), ETransStream(nullptr), potentialStream(nullptr), BStream(nullptr), AsqStream(nullptr),
      timeStream(nullptr), Ex0(0.0) {
    const size_t M = x_size + n_prepad + n_postpad;
    charge.assign(x_size, 0.0); J.assign(x_size, 0.0); neutralizationCharge.assign(x_size, 0.0);
    charges.assign(settings.q.size(), charge);
    for (std::vector<double>* y : {&By, &Bz, &Ey, &Ez, &Ay, &Az}) y->assign(8 * M, 0.0);
    a_squared.assign(x_size + 1, 0.0);
    PHI = new double[x_size]();
    fieldCoef = 1.0 / (12 * settings.GetDx(0));
    // the dense N x N Poisson matrix of the reference (EMSolver.cpp:28-86) does not exist here: vrt_poisson solves the same
    // linear system directly in O(N)
}
EMFieldSolver::~EMFieldSolver() {
    delete[] PHI;
    for (std::ofstream* s : {chargeStream, ELongStream, ETransStream, potentialStream, BStream, AsqStream, timeStream}) delete s;
}

void EMFieldSolver::AssembleRhoAndJ() { settings.Check(vrt_moments(settings.Gpu()), "vrt_moments"); mirrors_current_ = false; }
// EMFieldSolver::AssembleEnergy (EMSolver.cpp:124-131): dN/dp per species on the finest p grid
void EMFieldSolver::AssembleEnergy() {
    if (energies.empty()) for (unsigned int i = 0; i < meshes.size(); i++) energies.push_back(std::vector<double>(settings.p_size_finest[i], 0.0));
    for (auto& e : energies) std::fill(e.begin(), e.end(), 0.0);
    for (unsigned int i = 0; i < meshes.size(); i++) meshes[i]->InterpolateEnergyToFinestMesh(energies[i]);
}
void EMFieldSolver::UpdatePotential() { settings.Check(vrt_poisson(settings.Gpu()), "vrt_poisson"); mirrors_current_ = false; }
// EMFieldSolver::RGKStep (EMSolver.cpp:194-202); the laser inflow values are the user's GetBY/GetBZ at the current settings.time
void EMFieldSolver::RGKStep(int step, double timestep) {
    settings.Check(vrt_field_stage(settings.Gpu(), step, timestep, settings.GetBY(0, settings.time), settings.GetBZ(0, settings.time)), "vrt_field_stage");
    mirrors_current_ = false;
}
double EMFieldSolver::EstimateCFLBound() {
    double b = 0.0;
    settings.Check(vrt_cfl_bound(settings.Gpu(), &b), "vrt_cfl_bound");
    return b;
}
void EMFieldSolver::EnforceChargeNeutralization() {
    settings.Check(vrt_enforce_neutralization(settings.Gpu()), "vrt_enforce_neutralization");
    mirrors_current_ = false;
}

void EMFieldSolver::SyncHost() {
    if (mirrors_current_) return;
    vrt_ctx* g = settings.Gpu();
    const size_t M = x_size + n_prepad + n_postpad;
    std::vector<double>* Y[6] = {&By, &Bz, &Ey, &Ez, &Ay, &Az};
    for (int w = 0; w < 6; w++)
        for (int slot = 0; slot < 8; slot++) settings.Check(vrt_field_download(g, w, slot, Y[w]->data() + slot * M), "vrt_field_download");
    settings.Check(vrt_get_1d(g, VRT_CHARGE, charge.data()), "vrt_get_1d");
    settings.Check(vrt_get_1d(g, VRT_J, J.data()), "vrt_get_1d");
    settings.Check(vrt_get_1d(g, VRT_NEUTRALIZATION, neutralizationCharge.data()), "vrt_get_1d");
    settings.Check(vrt_get_1d(g, VRT_A_SQUARED, a_squared.data()), "vrt_get_1d");
    settings.Check(vrt_get_1d(g, VRT_PHI, PHI), "vrt_get_1d");
    for (size_t s = 0; s < charges.size(); s++) settings.Check(vrt_get_1d(g, VRT_CHARGES0 + (int)s, charges[s].data()), "vrt_get_1d");
    settings.Check(vrt_get_scalar(g, VRT_EX0, &Ex0), "vrt_get_scalar");
    mirrors_current_ = true;
}

double EMFieldSolver::GetASquared(int i) { SyncHost(); return a_squared[std::min(std::max(i, 0), (int)x_size)]; }
double EMFieldSolver::GetCellAverageASquared(int i) {
    SyncHost();
    i += n_prepad;
    i = std::min(std::max(i, 0), (int)x_size + n_prepad + n_postpad - 1);
    const double ay = Ay[Index(i, 1)], az = Az[Index(i, 1)];
    return (ay * ay) + (az * az);
}
// EMFieldSolver::GetMagneticForce (EMSolver.cpp:666-673): (A x B)_x of the stage values at plasma cell i
double EMFieldSolver::GetMagneticForce(int i) {
    SyncHost();
    i += n_prepad;
    i = std::min(std::max(i, 0), (int)x_size + n_prepad + n_postpad - 1);
    return Az[Index(i, 1)] * By[Index(i, 1)] - Ay[Index(i, 1)] * Bz[Index(i, 1)];
}
// EMFieldSolver::GetEfield (EMSolver.cpp:137-154) from the mirrored potential
double EMFieldSolver::GetEfield(int i) {
    SyncHost();
    const int N = (int)x_size;
    auto wrap = [N](int k) { k = k > -1 ? k : k + N; return k < N ? k : k - N; };
    return -fieldCoef * (8 * (PHI[wrap(i + 1)] - PHI[wrap(i - 1)]) - PHI[wrap(i + 2)] + PHI[wrap(i - 2)]) + Ex0;
}

namespace {
std::ofstream* open_dump(const char* name, int precision) {
    std::ofstream* s = new std::ofstream(name);
    (*s) << std::scientific << std::setprecision(precision);
    return s;
}
void write_row(std::ofstream& s, const double* v, size_t n) {
    for (size_t i = 0; i < n; i++) s << v[i] << " ";
    s << std::endl;
}
}  // namespace
// text dumps in the reference's formats (EMSolver.cpp:340-477): one line per array and call, values separated by blanks
void EMFieldSolver::DumpCharge() {
    SyncHost();
    if (!chargeStream) chargeStream = open_dump("output/charge.txt", settings.output.precision);
    for (auto& c : charges) write_row(*chargeStream, c.data(), c.size());
}
void EMFieldSolver::DumpEnergy() {                             // EMSolver.cpp:362-379: one file per species, one line per call
    if (energyStreams.empty())
        for (unsigned int i = 0; i < settings.q.size(); i++) {
            std::stringstream name;
            name << "output/dNdP_" << i << ".txt";
            energyStreams.push_back(std::make_unique<std::ofstream>(name.str()));
            (*energyStreams[i]) << std::scientific << std::setprecision(settings.output.precision);
        }
    for (unsigned int k = 0; k < settings.q.size() && k < energies.size(); k++) write_row(*energyStreams[k], energies[k].data(), energies[k].size());
}
void EMFieldSolver::DumpEFieldLongitudinal() {
    SyncHost();
    if (!ELongStream) ELongStream = open_dump("output/EFieldLong.txt", settings.output.precision);
    std::vector<double> e(x_size);
    for (unsigned int i = 0; i < x_size; i++) e[i] = GetEfield(i);
    write_row(*ELongStream, e.data(), e.size());
}
void EMFieldSolver::DumpEFieldTransverse() {
    SyncHost();
    if (!ETransStream) ETransStream = open_dump("output/EFieldTrans.txt", settings.output.precision);
    write_row(*ETransStream, Ey.data(), n_prepad + x_size + n_postpad);
    write_row(*ETransStream, Ez.data(), n_prepad + x_size + n_postpad);
}
void EMFieldSolver::DumpPotential() {
    SyncHost();
    if (!potentialStream) potentialStream = open_dump("output/potential.txt", settings.output.precision);
    write_row(*potentialStream, PHI, x_size);
}
void EMFieldSolver::DumpBFieldTransverse() {
    SyncHost();
    if (!BStream) BStream = open_dump("output/BFieldTrans.txt", settings.output.precision);
    write_row(*BStream, By.data(), n_prepad + x_size + n_postpad);
    write_row(*BStream, Bz.data(), n_prepad + x_size + n_postpad);
}
void EMFieldSolver::DumpAsqField() {
    SyncHost();
    if (!AsqStream) AsqStream = open_dump("output/ASquared.txt", settings.output.precision);
    write_row(*AsqStream, a_squared.data(), x_size);
}
void EMFieldSolver::DumpTime(double time) {
    if (!timeStream) timeStream = open_dump("output/time.txt", settings.output.precision);
    (*timeStream) << time << std::endl;
}

// =====================================================================================================================
// SolverManager
// =====================================================================================================================
SolverManager::SolverManager(Settings& settings) : settings(settings) {
    std::cout << "--- Time (T) ----|--- Particle ---|---- Levels ----|-------- Rectangles ---------" << std::endl;
    settings.Gpu();
    const int n_species = (int)settings.p_size.size();
    for (int s = 0; s < n_species; s++) meshes.push_back(std::make_shared<Mesh>(s, settings));
    EMSolver = std::make_shared<EMFieldSolver>(settings, meshes);
    std::cout << std::setfill(' ') << std::setw(4) << ' ' << std::fixed << 0.0 << std::setw(5) << ' ';
    for (auto& mesh : meshes) {
        mesh->SetFieldSolver(EMSolver);
        mesh->PushData();
        screenOutput(mesh);
    }
    EMSolver->EnforceChargeNeutralization();
}

// laser inflow values of the six stages: GetBY/GetBZ(0, time after UpdateTime(i)) (SolverManager.cpp:37-38, EMSolver.cpp:501-502)
void SolverManager::StageLasers(double timeStep, double laser[12]) {
    double t = settings.time;
    for (int i = 0; i < 6; i++) {
        t = vrt_update_time(t, i, timeStep);
        laser[2 * i] = settings.GetBY(0, t);
        laser[2 * i + 1] = settings.GetBZ(0, t);
    }
}
// SolverManager::Advance (SolverManager.cpp:28-39): the six stages replayed as one CUDA graph
void SolverManager::Advance(double timeStep) {
    double laser[12];
    StageLasers(timeStep, laser);
    settings.Check(vrt_set_scalar(settings.Gpu(), VRT_TIME, settings.time), "vrt_set_scalar");
    settings.Check(vrt_step(settings.Gpu(), timeStep, laser), "vrt_step");
    for (int i = 0; i < 6; i++) settings.UpdateTime(i, timeStep);
    for (auto& mesh : meshes) mesh->MarkDeviceCurrent();
    EMSolver->Invalidate();
}
void SolverManager::AdvanceFields(double timeStep) {      // SolverManager.cpp:41-46
    double laser[12];
    StageLasers(timeStep, laser);
    settings.Check(vrt_set_scalar(settings.Gpu(), VRT_TIME, settings.time), "vrt_set_scalar");
    settings.Check(vrt_step_fields(settings.Gpu(), timeStep, laser), "vrt_step_fields");
    for (int i = 0; i < 6; i++) settings.UpdateTime(i, timeStep);
    EMSolver->Invalidate();
}
void SolverManager::reGrid(double t) {                     // SolverManager.cpp:48-57
    const double T = settings.tempEM[0] / cs;
    std::cout << std::setfill(' ') << std::setw(4) << ' ' << std::fixed << std::setprecision(5) << t / T << std::setw(5) << ' ';
    for (auto& mesh : meshes) {
        mesh->updateHierarchy();
        screenOutput(mesh);
    }
    EMSolver->Invalidate();
}
double SolverManager::CalculateDt(double cfl) { return cfl * EMSolver->EstimateCFLBound(); }

std::string SolverManager::centeredOutput(std::string const& original, int targetSize) {
    const int padding = targetSize - (int)original.size();
    return padding > 0 ? std::string(padding / 2, ' ') + original + std::string(padding / 2, ' ') : original;
}
void SolverManager::screenOutput(const std::shared_ptr<Mesh>& mesh) {    // one table row: species, levels in use (of), rectangles per level
    std::stringstream counts, lv, pt;
    int used = 0;
    for (auto& lvl : mesh->levels)
        if (!lvl->rectangles.empty()) { used++; counts << lvl->rectangles.size() << ", "; }
    std::string cs_ = counts.str();
    if (cs_.size() >= 2) cs_.erase(cs_.size() - 2, 2);
    pt << mesh->particleType;
    lv << used << " (" << mesh->levels.size() << ")";
    if (mesh->particleType > 0) std::cout << std::setw(17) << ' ';
    std::cout << std::setw(16) << centeredOutput(pt.str(), 15) << ' ' << std::setw(16) << centeredOutput(lv.str(), 15) << ' '
              << std::setw(30) << centeredOutput(cs_, 29) << std::endl;
}
void SolverManager::OutputRectangles(double t) {
    if (!settings.output.rectangleData) return;
    for (auto& mesh : meshes) mesh->outputRectangleData(t);
}
void SolverManager::fileOutput(double t) {                 // SolverManager.cpp:102-160 (sequential: the mirrors are shared)
    const Output& o = settings.output;
    if (o.energy) { EMSolver->AssembleEnergy(); EMSolver->DumpEnergy(); }
    if (o.charge) EMSolver->DumpCharge();
    if (o.potential) EMSolver->DumpPotential();
    if (o.EFieldLongitudinal) EMSolver->DumpEFieldLongitudinal();
    if (o.EFieldTransverse) EMSolver->DumpEFieldTransverse();
    if (o.BFieldTransverse) EMSolver->DumpBFieldTransverse();
    if (o.AFieldSquared) EMSolver->DumpAsqField();
    if (o.time) EMSolver->DumpTime(t);
}
void SolverManager::Checkpoint(const std::string& path) {
    settings.Check(vrt_set_scalar(settings.Gpu(), VRT_TIME, settings.time), "vrt_set_scalar");
    settings.Check(vrt_checkpoint_write(settings.Gpu(), path.c_str()), "vrt_checkpoint_write");
}
void SolverManager::Restart(const std::string& path) {
    settings.Check(vrt_checkpoint_read(settings.Gpu(), path.c_str()), "vrt_checkpoint_read");
    double t = 0.0;
    settings.Check(vrt_get_scalar(settings.Gpu(), VRT_TIME, &t), "vrt_get_scalar");
    settings.time = t;
    for (auto& mesh : meshes) mesh->AdoptDeviceHierarchy();
    EMSolver->Invalidate();
}
void SolverManager::SyncHost() {
    for (auto& mesh : meshes) mesh->SyncHost();
    EMSolver->SyncHost();
}
