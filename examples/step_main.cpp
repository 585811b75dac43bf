// step_main.cpp -- a main program that keeps its OWN time loop, as examples/example1/src/Simulation.f90:58-115
// does over the library's Fortran subroutine API, written against the step-granular C entry points that the
// Fortran module RoadSurf of this repository forwards to (roadsurf_b200/fortran/RoadSurf.f90):
//
//     Initialization          -> roadsurf_session_open
//     SaveOutput(i)           -> roadsurf_step(session, i) + roadsurf_session_fetch(session, i, status)
//     CheckEndCoupling        -> the status word (RS_ST_FAILED ends the loop)
//     lastValues + last step  -> roadsurf_step(session, SimLen)
//
// One synthetic point, coupling + relaxation; the same point is then run through the reference's own entry
// `runsimulation` and the two result sets are compared value by value (they must be identical).  Without a
// CUDA device the program says so and exits 0 (the library has no CPU path).
//
//   g++ -std=c++17 -O2 -I include examples/step_main.cpp -o step_main -L roadsurf_b200 -lroadsurf_b200
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "roadsurf_b200.h"

int main(int argc, char** argv)
{
  const int chunk = argc > 1 ? std::atoi(argv[1]) : 1;  // 1 = strictly one launch per model step
  const int analysis_h = 2, forecast_h = 2;
  const double DT = 30.0;
  const int per_hour = static_cast<int>(3600.0 / DT);
  const int sim_len = 1 + (analysis_h + forecast_h) * per_hour;
  const int forecast_step = analysis_h * per_hour;

  InputSettings settings;
  roadsurf_default_settings(&settings, sim_len, DT);
  settings.use_coupling = 1;
  settings.use_relaxation = 1;
  settings.coupling_minutes = 45;
  InputParameters params;
  roadsurf_default_parameters(&params, DT);

  std::vector<int> year(sim_len, 2019), month(sim_len, 12), day(sim_len), hour(sim_len), minute(sim_len), second(sim_len);
  std::vector<double> tair(sim_len), tdew(sim_len), VZ(sim_len, 2.5), Rhz(sim_len, 90.0), prec(sim_len, 0.0), SW(sim_len, 0.0),
      LW(sim_len, 290.0), SW_dir(sim_len, 0.0), LW_net(sim_len, -25.0), obs(sim_len, -9999.9), depth(sim_len, -9999.9),
      horizons(360, 0.0);
  std::vector<int> phase(sim_len, -9999);
  for (int i = 0; i < sim_len; ++i)
  {
    const long t = 20L * 3600 + static_cast<long>(i * DT);
    day[i] = 1 + static_cast<int>(t / 86400);
    hour[i] = static_cast<int>((t % 86400) / 3600);
    minute[i] = static_cast<int>((t % 3600) / 60);
    second[i] = static_cast<int>(t % 60);
    const double h = i * DT / 3600.0;
    tair[i] = -0.5 + 2.0 * std::sin(2 * M_PI * (h - 5.0) / 24.0);
    tdew[i] = tair[i] - 1.0;
    if (i <= forecast_step) obs[i] = tair[i] - 0.6;
    if (h > 1.0 && h < 1.5) prec[i] = 0.8;
  }
  auto run = [&](bool stepwise, std::vector<double> (&o)[6], int* status) -> int {
    for (auto& v : o) v.assign(sim_len, 12345.0);
    std::vector<double> obs_copy = obs;  // read_input blanks the observations over the coupling window in place
    InputPointers ip;
    std::memset(&ip, 0, sizeof ip);
    ip.inputLen = sim_len;
    ip.c_tair = tair.data(); ip.c_tdew = tdew.data(); ip.c_VZ = VZ.data(); ip.c_Rhz = Rhz.data(); ip.c_prec = prec.data();
    ip.c_SW = SW.data(); ip.c_LW = LW.data(); ip.c_SW_dir = SW_dir.data(); ip.c_LW_net = LW_net.data();
    ip.c_TSurfObs = obs_copy.data(); ip.c_PrecPhase = phase.data(); ip.c_local_horizons = horizons.data();
    ip.c_Depth = depth.data(); ip.c_year = year.data(); ip.c_month = month.data(); ip.c_day = day.data();
    ip.c_hour = hour.data(); ip.c_minute = minute.data(); ip.c_second = second.data();
    OutputPointers op;
    op.outputLen = sim_len;
    op.c_TsurfOut = o[0].data(); op.c_SnowOut = o[1].data(); op.c_WaterOut = o[2].data();
    op.c_IceOut = o[3].data(); op.c_DepositOut = o[4].data(); op.c_Ice2Out = o[5].data();
    LocalParameters lp;
    std::memset(&lp, 0, sizeof lp);
    lp.tair_relax = lp.VZ_relax = lp.RH_relax = -9999.9;
    lp.couplingIndexI = -9999;
    lp.couplingTsurf = -9999.9;
    lp.lat = 60.2; lp.lon = 24.9; lp.sky_view = 1.0;
    const InputPointers* pin = &ip;
    OutputPointers* pout = &op;
    LocalParameters* ploc = &lp;
    const LocalParameters* pcloc = &lp;
    int latest = forecast_step + 1, ok = 0;
    if (roadsurf_read_input_derive(1, &pin, &settings, forecast_step, &latest, &ploc, &ok) != RS_OK || !ok) return 1;
    if (!stepwise)
    {
      runsimulation(&op, &ip, &settings, &params, &lp);
      *status = 0;
      return 0;
    }
    void* session = nullptr;                                              // ---- Initialization
    if (roadsurf_session_open(1, &pout, &pin, &settings, &params, &pcloc, &session) != RS_OK) return 2;
    roadsurf_session_set_chunk(session, chunk);
    int i = 1, rc = RS_OK;
    bool failed = false;
    while (i < sim_len && !failed && rc == RS_OK)                         // ---- Simulation.f90:58
    {
      rc = roadsurf_step(session, i);                                     // CheckValues ... CalcAlbedo, fused
      if (rc == RS_OK) rc = roadsurf_session_fetch(session, i, status);   // SaveOutput(i)
      failed = (*status & RS_ST_FAILED) != 0;                             // CheckEndCoupling: simulation_failed
      ++i;
    }
    if (!failed && rc == RS_OK)                                           // ---- Simulation.f90:100-115
    {
      rc = roadsurf_step(session, sim_len);
      if (rc == RS_OK) rc = roadsurf_session_fetch(session, sim_len, status);
    }
    roadsurf_session_close(session);
    return rc == RS_OK ? 0 : 3;
  };

  std::printf("%s: 1 point x %d steps, run-ahead chunk %d\n", roadsurf_version(), sim_len, chunk);
  if (roadsurf_device_count() < 1)
  {
    std::printf("no CUDA device visible: nothing was run (the library has no CPU path)\n");
    return 0;
  }
  std::vector<double> a[6], b[6];
  int st_a = 0, st_b = 0;
  if (int rc = run(true, a, &st_a))
  {
    std::fprintf(stderr, "stepwise run failed (%d): %s\n", rc, roadsurf_last_error());
    return 1;
  }
  RsLaunchInfo li;
  roadsurf_last_launch(&li);
  const int launches_stepwise = li.launches_total;
  if (int rc = run(false, b, &st_b))
  {
    std::fprintf(stderr, "runsimulation failed (%d): %s\n", rc, roadsurf_last_error());
    return 1;
  }
  long differing = 0;
  for (int v = 0; v < 6; ++v)
    for (int t = 0; t < sim_len; ++t)
      differing += std::memcmp(&a[v][t], &b[v][t], sizeof(double)) != 0;
  std::printf("kernel launches of the stepwise run: %d; status %d; Tsurf(end) %.6f; values differing from runsimulation: %ld\n",
              launches_stepwise, st_a, a[0][sim_len - 1], differing);
  return differing == 0 ? 0 : 2;
}
