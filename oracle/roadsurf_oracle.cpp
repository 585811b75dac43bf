// roadsurf_oracle.cpp -- C exports of the CPU oracle (see roadsurf_oracle.hpp header comment).
//
// TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference golden vectors exist, no Fortran
// compiler here).  Built by oracle/Makefile into oracle/_build/:
//   liboracle.so       g++ -O2 -ffp-contract=off            parity oracle (no FMA, no fast-math)
//   liboracle_fast.so  g++ -O2 -Ofast + the reference's MATHFLAGS (Makefile:28,37 of the
//                      reference) -- the timing baseline "compiled like the reference"
#include "roadsurf_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace rs_oracle
{
// ---- operation-counting scalar ------------------------------------------------------------------
struct OpCounts
{
  unsigned long long add = 0, mul = 0, div = 0, sqrt = 0, exp = 0, log = 0, trig = 0, pow = 0,
                     cmp = 0;
};
static thread_local OpCounts g_ops;

struct Counted
{
  double v;
  Counted() : v(0.0) {}
  Counted(double x) : v(x) {}
  Counted(int x) : v(x) {}
  Counted(long x) : v(static_cast<double>(x)) {}
};
inline Counted operator+(Counted a, Counted b) { ++g_ops.add; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { ++g_ops.add; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { ++g_ops.mul; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { ++g_ops.div; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline bool operator<(Counted a, Counted b) { ++g_ops.cmp; return a.v < b.v; }
inline bool operator>(Counted a, Counted b) { ++g_ops.cmp; return a.v > b.v; }
inline bool operator<=(Counted a, Counted b) { ++g_ops.cmp; return a.v <= b.v; }
inline bool operator>=(Counted a, Counted b) { ++g_ops.cmp; return a.v >= b.v; }
inline Counted r_sqrt(Counted x) { ++g_ops.sqrt; return Counted(std::sqrt(x.v)); }
inline Counted r_exp(Counted x) { ++g_ops.exp; return Counted(std::exp(x.v)); }
inline Counted r_log(Counted x) { ++g_ops.log; return Counted(std::log(x.v)); }
inline Counted r_sin(Counted x) { ++g_ops.trig; return Counted(std::sin(x.v)); }
inline Counted r_cos(Counted x) { ++g_ops.trig; return Counted(std::cos(x.v)); }
inline Counted r_acos(Counted x) { ++g_ops.trig; return Counted(std::acos(x.v)); }
inline Counted r_asin(Counted x) { ++g_ops.trig; return Counted(std::asin(x.v)); }
inline Counted r_atan2(Counted y, Counted x) { ++g_ops.trig; return Counted(std::atan2(y.v, x.v)); }
inline Counted r_pow(Counted x, Counted y) { ++g_ops.pow; return Counted(std::pow(x.v, y.v)); }
inline Counted r_abs(Counted x) { return Counted(std::fabs(x.v)); }
inline Counted r_aint(Counted x) { return Counted(std::trunc(x.v)); }
inline double r_val(Counted x) { return x.v; }
}  // namespace rs_oracle

using namespace rs_oracle;

namespace
{
template <class R>
int status_word(const Model<R>& m)
{
  int st = 0;
  if (m.settings.simulation_failed) st |= RS_ST_FAILED;
  if (m.diag.bad_input) st |= RS_ST_BAD_INPUT;
  if (m.diag.abnormal_tsurf) st |= RS_ST_ABNORMAL_TSURF;
  if (m.diag.coupling_used) st |= RS_ST_COUPLING_USED;
  if (m.diag.coupling_used && m.diag.coupling_failed) st |= RS_ST_COUPLING_FAILED;
  if (m.diag.bl_not_converged) st |= RS_ST_BL_NOT_CONVERGED;
  if (m.diag.solar_stop) st |= RS_ST_SOLAR_GEOMETRY;
  return st;
}
}  // namespace

extern "C" {

// Same signature and semantics as the reference's runsimulation
// (examples/example1/src/Simulation.f90:4-117), including the in-place mutation of the inputs.
void oracle_runsimulation(OutputPointers* out, const InputPointers* in, const InputSettings* settings,
                          const InputParameters* params, const LocalParameters* local)
{
  Model<double> m;
  m.runsimulation(*out, *in, *settings, *params, *local);
}

// As above, additionally returning the status word (include/roadsurf_b200.h RS_ST_*), the number
// of executed point-steps (coupling re-runs included) and boundary-layer iterations.
int oracle_runsimulation_ex(OutputPointers* out, const InputPointers* in, const InputSettings* settings,
                            const InputParameters* params, const LocalParameters* local,
                            long long* executed_steps, long long* bl_iterations, int verbose)
{
  Model<double> m;
  m.diag.verbose = verbose != 0;
  m.runsimulation(*out, *in, *settings, *params, *local);
  if (executed_steps) *executed_steps = m.diag.executed_steps;
  if (bl_iterations) *bl_iterations = m.diag.bl_iterations;
  return status_word(m);
}

// One point with a per-step trace of internals: trace[SimLen][16] (see roadModelOneStep).
int oracle_trace_point(OutputPointers* out, const InputPointers* in, const InputSettings* settings,
                       const InputParameters* params, const LocalParameters* local, double* trace)
{
  Model<double> m;
  m.diag.trace = trace;
  m.runsimulation(*out, *in, *settings, *params, *local);
  return status_word(m);
}

// Work-queue over points on `nthreads` host threads: the equivalent of example1's `-j N`
// (examples/example1/src/WorkQueue.h:15-130, roadrunner.cpp:423-501).  Used for the CPU baseline.
void oracle_run_batch(int npoints, OutputPointers* const* out, const InputPointers* const* in,
                      const InputSettings* settings, const InputParameters* params,
                      const LocalParameters* const* local, int nthreads, int* status,
                      long long* executed_steps)
{
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  std::atomic<long long> steps(0);
  auto worker = [&]() {
    long long mine = 0;
    for (;;)
    {
      const int p = next.fetch_add(1);
      if (p >= npoints) break;
      Model<double> m;
      m.runsimulation(*out[p], *in[p], *settings, *params, *local[p]);
      if (status) status[p] = status_word(m);
      mine += m.diag.executed_steps;
    }
    steps += mine;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  if (executed_steps) *executed_steps = steps.load();
}

// oracle_run_batch that also returns every point's final ground temperature profile and surface state:
// tmp_out[p][0..N+1] = ground%Tmp(0:N+1) after the last step, surf_out[p][0..9] = TsurfAve, SrfWatmms,
// SrfSnowmms, SrfIcemms, SrfIce2mms, SrfDepmms, Q2Melt, T4Melt, EvapmmTS, Albedo (the order of the
// kernel's state planes).  For the parity tests of the ground-layer temperatures.
void oracle_run_batch_state(int npoints, OutputPointers* const* out, const InputPointers* const* in,
                            const InputSettings* settings, const InputParameters* params,
                            const LocalParameters* const* local, int nthreads, int* status, double* tmp_out,
                            double* surf_out)
{
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;)
    {
      const int p = next.fetch_add(1);
      if (p >= npoints) break;
      Model<double> m;
      m.runsimulation(*out[p], *in[p], *settings, *params, *local[p]);
      if (status) status[p] = status_word(m);
      const int nl = m.settings.NLayers;
      if (tmp_out)
        for (int j = 0; j <= nl + 1; ++j) tmp_out[static_cast<size_t>(p) * (nl + 2) + j] = m.ground.Tmp[j];
      if (surf_out)
      {
        double* o = surf_out + static_cast<size_t>(p) * 10;
        o[0] = m.surf.TsurfAve;
        o[1] = m.surf.SrfWatmms;
        o[2] = m.surf.SrfSnowmms;
        o[3] = m.surf.SrfIcemms;
        o[4] = m.surf.SrfIce2mms;
        o[5] = m.surf.SrfDepmms;
        o[6] = m.surf.Q2Melt;
        o[7] = m.surf.T4Melt;
        o[8] = m.surf.EvapmmTS;
        o[9] = m.ground.Albedo;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
}

// Run one point with the op-counting scalar.  counts[9] = add, mul, div, sqrt, exp, log, trig,
// pow, cmp; returns the number of executed steps.
long long oracle_count_ops(OutputPointers* out, const InputPointers* in, const InputSettings* settings,
                           const InputParameters* params, const LocalParameters* local,
                           unsigned long long* counts)
{
  g_ops = OpCounts();
  Model<Counted> m;
  m.runsimulation(*out, *in, *settings, *params, *local);
  counts[0] = g_ops.add;
  counts[1] = g_ops.mul;
  counts[2] = g_ops.div;
  counts[3] = g_ops.sqrt;
  counts[4] = g_ops.exp;
  counts[5] = g_ops.log;
  counts[6] = g_ops.trig;
  counts[7] = g_ops.pow;
  counts[8] = g_ops.cmp;
  return m.diag.executed_steps;
}

// ---- example2's time interpolation of raw records (examples/example2/src/AsciiSource.cpp) --------

// AsciiSource::Impl::interpolate (:223-281) + the per-time loop of GetWeather (:292-345) for ONE variable
// of one point: raw[nrec] at model steps record_step[nrec] (times = step * DT seconds; the reference's
// NFmiMetTime::DifferenceInMinutes is taken as whole minutes) -> out[sim_len], left at `fill` where the
// reference assigns nothing.  kind 0: plain; 1: relative humidity (clamped to [0, 100], :322-323);
// 2: precipitation (> 100 dropped, :326-327); 3: precipitation phase (interpolated as a double, then
// stored into the int array, :345: truncation).  A value is missing when NaN or < -9000 (read_file
// :150-165 leaves MISSING = NaN where the file has -9999).
void oracle_interpolate_example2(const double* raw, const int* record_step, int nrec, double DT, int sim_len, int kind,
                                 double fill, double* out)
{
  auto is_missing = [](double x) { return std::isnan(x) || x < -9000.0; };
  auto minutes = [&](int steps) { return static_cast<long long>(static_cast<double>(steps) * DT) / 60; };
  const double MISSING = std::nan("");
  for (int i = 0; i < sim_len; ++i)
  {
    out[i] = fill;
    // :307-308 not in the time interval of the data
    if (nrec < 1 || i < record_step[0] || i > record_step[nrec - 1]) continue;
    // :313-316 first position where t >= time
    int pos = 0;
    while (pos < nrec && record_step[pos] < i) ++pos;
    double value = MISSING;
    bool done = false;
    // :228-233
    if (record_step[pos] == i && !is_missing(raw[pos]))
    {
      value = raw[pos];
      done = true;
    }
    // :235-236
    if (!done && pos == 0) done = true;
    if (!done)
    {
      // :240-250 first valid value at or after
      int pos2;
      double value2 = MISSING;
      for (pos2 = pos; pos2 < nrec; ++pos2)
      {
        value2 = raw[pos2];
        if (!is_missing(value2)) break;
      }
      if (!is_missing(value2))
      {
        // :254-264 first valid value before
        int pos1;
        double value1 = MISSING;
        for (pos1 = pos - 1;; --pos1)
        {
          value1 = raw[pos1];
          if (pos1 == 0 || !is_missing(value1)) break;
        }
        if (!is_missing(value1))
        {
          // :268-279
          const long long gap = minutes(record_step[pos2] - record_step[pos1]);
          if (gap <= 180 && gap > 0)
          {
            const long long gap1 = minutes(i - record_step[pos1]);
            value = (static_cast<double>(gap - gap1) * value1 + static_cast<double>(gap1) * value2) / static_cast<double>(gap);
          }
        }
      }
    }
    if (kind == 1 && !is_missing(value)) value = std::max(0.0, std::min(100.0, value));
    if (kind == 2 && value > 100) value = MISSING;
    if (!is_missing(value)) out[i] = (kind == 3) ? static_cast<double>(static_cast<int>(value)) : value;
  }
}

// ---- unit-level entry points for the known-answer tests ---------------------------------------

// src/Initialization.f90:217-235: ZDpth(1..N+1) -> z[0..N]
void oracle_layer_depths(int nlayers, double* z)
{
  Model<double> m;
  m.settings.NLayers = nlayers;
  m.allocator();
  m.initDepth();
  for (int i = 1; i <= nlayers + 1; ++i) z[i - 1] = m.ground.ZDpth[i];
}

// Per-run ground constants for a parameter set: condDZ, DyC, Wcont, CC (each [N], layer 1 first).
void oracle_ground_constants(const InputSettings* settings, const InputParameters* params, double* condDZ,
                             double* DyC, double* Wcont, double* CC)
{
  Model<double> m;
  LocalParameters lp;
  std::memset(&lp, 0, sizeof lp);
  m.initSettings(*settings, *params, lp);
  m.allocator();
  m.initDepth();
  m.InitParam(*params);
  m.HCapValues();
  m.ground_prop_init();
  m.CalcCC();
  for (int j = 1; j <= m.settings.NLayers; ++j)
  {
    condDZ[j - 1] = -(m.ground.CC[j] / m.ground.DyK[j]);
    DyC[j - 1] = m.ground.DyC[j];
    Wcont[j - 1] = m.ground.Wcont[j];
    CC[j - 1] = m.ground.CC[j];
  }
}

// src/BalanceModel.f90:325-351
int oracle_julday(int year, int month, int day)
{
  Model<double> m;
  m.modelInput.year = &year;
  m.modelInput.month = &month;
  m.modelInput.day = &day;
  return m.JulDay(1);
}

// src/SunPosition.f90:196-260
double oracle_jde(int year, int month, int day, int hour, int minute, int second)
{
  Model<double> m;
  m.modelInput.year = &year;
  m.modelInput.month = &month;
  m.modelInput.day = &day;
  m.modelInput.hour = &hour;
  m.modelInput.minute = &minute;
  m.modelInput.second = &second;
  return m.JulianEphemerisDay(1);
}

// src/SunPosition.f90:4-194; returns 1 if the reference would `stop`
int oracle_sun_position(int year, int month, int day, int hour, int minute, int second, double lat,
                        double lon, double* elevation, double* azimuth)
{
  Model<double> m;
  m.modelInput.year = &year;
  m.modelInput.month = &month;
  m.modelInput.day = &day;
  m.modelInput.hour = &hour;
  m.modelInput.minute = &minute;
  m.modelInput.second = &second;
  LocalParameters lp;
  std::memset(&lp, 0, sizeof lp);
  lp.lat = lat;
  lp.lon = lon;
  m.SunPosition(lp, 1, *elevation, *azimuth);
  return m.diag.solar_stop ? 1 : 0;
}

// oracle_sun_position for every (step, point) pair on `nthreads` threads: tf = [6][n_steps] time fields,
// elevation / azimuth [n_steps][npoints]; NaN where the reference would `stop`.
void oracle_sun_position_batch(const int* tf, int n_steps, const double* lat, const double* lon, int npoints, int nthreads,
                               double* elevation, double* azimuth)
{
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto worker = [&]() {
    for (;;)
    {
      const int t = next.fetch_add(1);
      if (t >= n_steps) break;
      for (int p = 0; p < npoints; ++p)
      {
        double e, a;
        const int stop = oracle_sun_position(tf[t], tf[n_steps + t], tf[2 * n_steps + t], tf[3 * n_steps + t],
                                             tf[4 * n_steps + t], tf[5 * n_steps + t], lat[p], lon[p], &e, &a);
        elevation[static_cast<size_t>(t) * npoints + p] = stop ? std::nan("") : e;
        azimuth[static_cast<size_t>(t) * npoints + p] = stop ? std::nan("") : a;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int k = 1; k < nthreads; ++k) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
}

// src/InputOutput.f90:239-268 and :202-236
double oracle_calc_tdew(double t2m, double rh) { return Model<double>::CalcTDew(t2m, rh); }
double oracle_calc_rh(double t2m, double tdew) { return Model<double>::CalcRhOne(t2m, tdew); }

// src/Cond.f90:143-249: returns PrecType; rain/snow in mm per time step
int oracle_prec_type(const InputSettings* settings, const InputParameters* params, int phase,
                     double prec_mm_h, double tair, double rh, double* rain, double* snow)
{
  Model<double> m;
  LocalParameters lp;
  std::memset(&lp, 0, sizeof lp);
  m.initSettings(*settings, *params, lp);
  m.condInit(*params);
  m.atm.Tair = tair;
  m.atm.RHz = rh;
  m.atm.PrecInTStep = prec_mm_h / 3600 * m.settings.DTSecs;
  m.atm.SnowType = SURFACE_SNOW_DRY;
  m.CalcPrecType(phase);
  *rain = m.atm.RainmmTS;
  *snow = m.atm.SnowmmTS;
  return m.atm.PrecType;
}

// src/BoundaryLayer.f90:3-190.  io = {BLCond, LE_Flux, EvapmmTS}; returns the iteration count.
int oracle_boundary_layer(const InputSettings* settings, const InputParameters* params, double tair,
                          double vz, double rh, double tsurf, double water, double* io)
{
  Model<double> m;
  LocalParameters lp;
  std::memset(&lp, 0, sizeof lp);
  m.initSettings(*settings, *params, lp);
  m.InitParam(*params);
  m.atm.Tair = tair;
  m.atm.VZ = vz;
  m.atm.RHz = rh;
  m.atm.BLCond = -99.9;
  m.surf.TsurfAve = tsurf;
  m.surf.SrfWatmms = water;
  m.surf.EvapmmTS = 0.0;
  m.CalcBLCondAndLE();
  io[0] = m.atm.BLCond;
  io[1] = m.atm.LE_Flux;
  io[2] = m.surf.EvapmmTS;
  return static_cast<int>(m.diag.bl_iterations);
}

// One call of WearFactors + RoadCond + CalcAlbedo (src/Cond.f90:9-139) on explicit state.
// st = {TsurfAve, Wat, Snow, Ice, Ice2, Dep, Q2Melt, T4Melt, EvapmmTS, Albedo} in and out.
void oracle_road_cond(const InputSettings* settings, const InputParameters* params, double* st)
{
  Model<double> m;
  LocalParameters lp;
  std::memset(&lp, 0, sizeof lp);
  m.initSettings(*settings, *params, lp);
  m.settings.Tph = m.settings.DTSecs / 3600.0;
  m.InitParam(*params);
  m.initSurf(true);
  m.condInit(*params);
  m.surf.TsurfAve = st[0];
  m.surf.SrfWatmms = st[1];
  m.surf.SrfSnowmms = st[2];
  m.surf.SrfIcemms = st[3];
  m.surf.SrfIce2mms = st[4];
  m.surf.SrfDepmms = st[5];
  m.surf.Q2Melt = st[6];
  m.surf.T4Melt = st[7];
  m.surf.EvapmmTS = st[8];
  m.ground.Albedo = st[9];
  m.atm.SnowType = SURFACE_SNOW_DRY;
  WearingFactors<double> w;
  m.WearFactors(w);
  m.RoadCond(m.phy.MaxPormms, w);
  m.CalcAlbedo();
  st[0] = m.surf.TsurfAve;
  st[1] = m.surf.SrfWatmms;
  st[2] = m.surf.SrfSnowmms;
  st[3] = m.surf.SrfIcemms;
  st[4] = m.surf.SrfIce2mms;
  st[5] = m.surf.SrfDepmms;
  st[6] = m.surf.Q2Melt;
  st[7] = m.surf.T4Melt;
  st[8] = m.surf.EvapmmTS;
  st[9] = m.ground.Albedo;
}

// src/Coupling.f90:292-481 on explicit state.
// c = {TsurfAve, lastTsurfObs, RadCoeff, RadCoeffPrevious, TsurfNearestAbove, TsurfNearestBelow,
//      RadCoefNearestAbove, RadCoefNearestBelow, SWRadCof, LWRadCof, SW_correction, LW_correction,
//      Tsurf_end_coup1}; flags = {Coupling_iterations, Coupling_failed, start_coupling_again}.
void oracle_coupling_control(double* c, int* flags)
{
  Model<double> m;
  CouplingVariables<double>& k = m.coupling;
  k.lastTsurfObs = c[1];
  k.RadCoeff = c[2];
  k.RadCoeffPrevious = c[3];
  k.TsurfNearestAbove = c[4];
  k.TsurfNearestBelow = c[5];
  k.RadCoefNearestAbove = c[6];
  k.RadCoefNearestBelow = c[7];
  k.SWRadCof = c[8];
  k.LWRadCof = c[9];
  k.SW_correction = c[10];
  k.LW_correction = c[11];
  k.Tsurf_end_coup1 = c[12];
  k.Coupling_iterations = flags[0];
  k.Coupling_failed = flags[1] != 0;
  k.start_coupling_again = flags[2] != 0;
  k.CoupPhaseN = 1;
  k.NObs = 1;
  double ts = c[0];
  m.CouplingOperations2(ts);
  c[0] = ts;
  c[1] = k.lastTsurfObs;
  c[2] = k.RadCoeff;
  c[3] = k.RadCoeffPrevious;
  c[4] = k.TsurfNearestAbove;
  c[5] = k.TsurfNearestBelow;
  c[6] = k.RadCoefNearestAbove;
  c[7] = k.RadCoefNearestBelow;
  c[8] = k.SWRadCof;
  c[9] = k.LWRadCof;
  c[10] = k.SW_correction;
  c[11] = k.LW_correction;
  c[12] = k.Tsurf_end_coup1;
  flags[0] = k.Coupling_iterations;
  flags[1] = k.Coupling_failed ? 1 : 0;
  flags[2] = k.start_coupling_again ? 1 : 0;
}

const char* oracle_build_flavour(void)
{
#ifdef RS_ORACLE_FAST
  return "fast (-O2 -Ofast + reference MATHFLAGS)";
#else
  return "parity (-O2 -ffp-contract=off)";
#endif
}

}  // extern "C"
