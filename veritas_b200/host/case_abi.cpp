// C-ABI access to the laser-plasma case functions (laser_plasma_case.hpp) so that non-C++ hosts
// (the Python tests and bench.py) evaluate the user-level case with the same arithmetic as the
// C++ host classes.  Restates /root/reference/veritas.cpp:36-115 (see the header).
#include "laser_plasma_case.hpp"
#include "../../include/veritas_b200.h"

extern "C" {

int vrt_case_derive(const vrt_case_params* p, double m0, double q0, unsigned x_size, const unsigned* p_size,
                    double refinementCriteria, vrt_case_derived* out) {
    if (!p || !out || !p_size) return -1;
    vrt_case::LaserPlasma c;
    c.lambda = p->lambda; c.a0 = p->a0; c.density = p->density; c.temp_frac = p->temp_frac;
    c.pmax_e = p->pmax_e; c.pmax_i = p->pmax_i; c.box_lambdas = p->box_lambdas; c.ion_mass_ratio = p->ion_mass_ratio;
    vrt_case::Derived d = vrt_case::derive(c, m0, q0, x_size, p_size, refinementCriteria);
    for (int s = 0; s < 2; s++) { out->dp[s] = d.dp[s]; out->pmin[s] = d.pmin[s]; out->temp0[s] = d.temp0[s]; out->temp1[s] = d.temp1[s]; out->tempEM[s] = d.tempEM[s]; }
    out->dx = d.dx; out->sizeWeight = d.sizeWeight; out->quadratureDepth = d.quadratureDepth;
    return 0;
}
double vrt_case_laser_by(double lambda, double amp, double x, double t) { return vrt_case::laser_by(lambda, amp, x, t); }
double vrt_case_laser_bz(double lambda, double amp, double x, double t) { return vrt_case::laser_bz(lambda, amp, x, t); }
double vrt_case_maxwellian_slab(double x, double p, double xl, double xr, double n0, double T) {
    return vrt_case::maxwellian_slab(x, p, xl, xr, n0, T);
}

}  // extern "C"
