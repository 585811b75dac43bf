// batch_main.cpp -- a C++ main program over the C ABI, shaped like the reference's example1
// (examples/example1/src/roadrunner.cpp:423-501: one InputData / OutputData / LocalParameters per
// point, then one model run per point) but handing ALL points to roadsurf_run_batch at once.
//
// It builds a small synthetic forecast (no files), derives the per-point parameters with
// roadsurf_read_input_derive exactly where the reference calls read_input, runs the batch and prints
// a few values.  Without a CUDA device it reports that and exits 0 (the library has no CPU path).
//
// With a second argument T > 0 the same points then go through the reference's own entry `runsimulation`,
// one point per call from a pool of T threads -- the shape of the reference's main
// (examples/example1/src/roadrunner.cpp:454-496) -- and the two result sets are compared value by value.
//
//   g++ -std=c++17 -O2 -pthread -I include examples/batch_main.cpp -o batch_main -L roadsurf_b200 -lroadsurf_b200
//   (plus -Wl,-rpath,<repo>/roadsurf_b200 or LD_LIBRARY_PATH)
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "roadsurf_b200.h"

namespace
{
struct PointData  // the vectors the reference keeps in InputData / OutputData (InputData.cpp:28-50)
{
  std::vector<double> tair, tdew, VZ, Rhz, prec, SW, LW, SW_dir, LW_net, TSurfObs, depth, horizons;
  std::vector<int> phase;
  std::vector<double> Tsurf, Snow, Water, Ice, Deposit, Ice2;
};
}  // namespace

int main(int argc, char** argv)
{
  const int npoints = argc > 1 ? std::atoi(argv[1]) : 64;
  const int pool_threads = argc > 2 ? std::atoi(argv[2]) : 0;
  const int analysis_h = 3, forecast_h = 6;
  const double DT = 30.0;
  const int per_hour = static_cast<int>(3600.0 / DT);
  const int sim_len = 1 + (analysis_h + forecast_h) * per_hour;
  const int forecast_step = analysis_h * per_hour;

  InputSettings settings;
  roadsurf_default_settings(&settings, sim_len, DT);  // (also zeroes force_tsurf, which the C++ examples leave unset)
  settings.use_coupling = 1;
  settings.use_relaxation = 1;
  settings.coupling_minutes = 60;

  InputParameters params;
  roadsurf_default_parameters(&params, DT);

  // shared time axis: 2019-12-01 21:00 UTC + i * DT
  std::vector<int> year(sim_len, 2019), month(sim_len, 12), day(sim_len), hour(sim_len), minute(sim_len), second(sim_len);
  for (int i = 0; i < sim_len; ++i)
  {
    const long t = 21L * 3600 + static_cast<long>(i * DT);
    day[i] = 1 + static_cast<int>(t / 86400);
    hour[i] = static_cast<int>((t % 86400) / 3600);
    minute[i] = static_cast<int>((t % 3600) / 60);
    second[i] = static_cast<int>(t % 60);
  }

  std::vector<PointData> pts(npoints);
  std::vector<InputPointers> in(npoints);
  std::vector<OutputPointers> out(npoints);
  std::vector<LocalParameters> local(npoints);
  for (int p = 0; p < npoints; ++p)
  {
    PointData& d = pts[p];
    auto fill = [&](std::vector<double>& v, double x) { v.assign(sim_len, x); };
    fill(d.tair, 0.0); fill(d.tdew, 0.0); fill(d.VZ, 3.0); fill(d.Rhz, 85.0); fill(d.prec, 0.0);
    fill(d.SW, 0.0); fill(d.LW, 280.0); fill(d.SW_dir, 0.0); fill(d.LW_net, -30.0);
    fill(d.TSurfObs, -9999.9); fill(d.depth, -9999.9);
    d.phase.assign(sim_len, -9999);
    d.horizons.assign(360, 0.0);
    for (int i = 0; i < sim_len; ++i)
    {
      const double h = i * DT / 3600.0;
      d.tair[i] = -1.0 + 0.05 * p + 3.0 * std::sin(2 * M_PI * (h - 6.0) / 24.0);
      d.tdew[i] = d.tair[i] - 1.5;
      if (i <= forecast_step) d.TSurfObs[i] = d.tair[i] - 0.8;  // road observations during the analysis
      if (h > 4.0 && h < 5.0) d.prec[i] = 0.6;                   // an hour of precipitation, phase to be interpreted
    }
    for (auto* v : {&d.Tsurf, &d.Snow, &d.Water, &d.Ice, &d.Deposit, &d.Ice2}) v->assign(sim_len, -9999.0);
    InputPointers& ip = in[p];
    std::memset(&ip, 0, sizeof ip);
    ip.inputLen = sim_len;
    ip.c_tair = d.tair.data(); ip.c_tdew = d.tdew.data(); ip.c_VZ = d.VZ.data(); ip.c_Rhz = d.Rhz.data();
    ip.c_prec = d.prec.data(); ip.c_SW = d.SW.data(); ip.c_LW = d.LW.data(); ip.c_SW_dir = d.SW_dir.data();
    ip.c_LW_net = d.LW_net.data(); ip.c_TSurfObs = d.TSurfObs.data(); ip.c_PrecPhase = d.phase.data();
    ip.c_local_horizons = d.horizons.data(); ip.c_Depth = d.depth.data();
    ip.c_year = year.data(); ip.c_month = month.data(); ip.c_day = day.data(); ip.c_hour = hour.data();
    ip.c_minute = minute.data(); ip.c_second = second.data();
    OutputPointers& op = out[p];
    op.outputLen = sim_len;
    op.c_TsurfOut = d.Tsurf.data(); op.c_SnowOut = d.Snow.data(); op.c_WaterOut = d.Water.data();
    op.c_IceOut = d.Ice.data(); op.c_DepositOut = d.Deposit.data(); op.c_Ice2Out = d.Ice2.data();
    LocalParameters& lp = local[p];
    std::memset(&lp, 0, sizeof lp);
    lp.tair_relax = lp.VZ_relax = lp.RH_relax = -9999.9;
    lp.couplingIndexI = -9999;
    lp.couplingTsurf = -9999.9;
    lp.lat = 60.2 + 0.01 * p;
    lp.lon = 24.9;
    lp.sky_view = 1.0;
  }

  std::vector<const InputPointers*> in_ptr(npoints);
  std::vector<OutputPointers*> out_ptr(npoints);
  std::vector<LocalParameters*> loc_ptr(npoints);
  std::vector<const LocalParameters*> loc_cptr(npoints);
  for (int p = 0; p < npoints; ++p)
  {
    in_ptr[p] = &in[p]; out_ptr[p] = &out[p]; loc_ptr[p] = &local[p]; loc_cptr[p] = &local[p];
  }
  // what read_input derives (roadrunner.cpp:157-278): InitLenI, relaxation targets, coupling index / obs
  std::vector<int> latest_obs(npoints, forecast_step + 1), ok(npoints, 0);
  if (roadsurf_read_input_derive(npoints, in_ptr.data(), &settings, forecast_step, latest_obs.data(), loc_ptr.data(),
                                 ok.data()) != RS_OK)
  {
    std::fprintf(stderr, "derive failed: %s\n", roadsurf_last_error());
    return 1;
  }
  std::printf("%s: %d points x %d steps, coupling window ends at index %d\n", roadsurf_version(), npoints,
              sim_len, local[0].couplingIndexI);
  if (roadsurf_device_count() < 1)
  {
    std::printf("no CUDA device visible: nothing was run (the library has no CPU path)\n");
    return 0;
  }
  std::vector<int> status(npoints, 0);
  using clock = std::chrono::steady_clock;
  auto ms_since = [](clock::time_point t0) { return std::chrono::duration<double, std::milli>(clock::now() - t0).count(); };
  roadsurf_device_count();
  auto t0 = clock::now();
  const int rc = roadsurf_run_batch(npoints, out_ptr.data(), in_ptr.data(), &settings, &params, loc_cptr.data(), 1,
                                    status.data());
  if (rc != RS_OK)
  {
    std::fprintf(stderr, "run failed (%d): %s\n", rc, roadsurf_last_error());
    return 1;
  }
  const double batch_ms = ms_since(t0);
  if (pool_threads > 0)
  {
    // the reference's way: a pool of host threads, each taking the next point and calling runsimulation on it
    std::vector<std::vector<double>> o2(static_cast<size_t>(npoints) * 6, std::vector<double>(sim_len, -9999.0));
    std::vector<OutputPointers> out2(npoints);
    for (int p = 0; p < npoints; ++p)
    {
      out2[p].outputLen = sim_len;
      out2[p].c_TsurfOut = o2[6 * p + 0].data(); out2[p].c_SnowOut = o2[6 * p + 1].data();
      out2[p].c_WaterOut = o2[6 * p + 2].data(); out2[p].c_IceOut = o2[6 * p + 3].data();
      out2[p].c_DepositOut = o2[6 * p + 4].data(); out2[p].c_Ice2Out = o2[6 * p + 5].data();
    }
    long long calls0 = 0, batches0 = 0, calls1 = 0, batches1 = 0;
    roadsurf_runsimulation_counters(&calls0, &batches0);
    std::atomic<int> next{0};
    t0 = clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < pool_threads; ++t)
      pool.emplace_back([&] {
        for (int p = next++; p < npoints; p = next++) runsimulation(&out2[p], &in[p], &settings, &params, &local[p]);
      });
    for (auto& th : pool) th.join();
    const double pool_ms = ms_since(t0);
    roadsurf_runsimulation_counters(&calls1, &batches1);
    long differing = 0;
    for (int p = 0; p < npoints; ++p)
    {
      const std::vector<double>* a[6] = {&pts[p].Tsurf, &pts[p].Snow, &pts[p].Water, &pts[p].Ice, &pts[p].Deposit, &pts[p].Ice2};
      for (int v = 0; v < 6; ++v) differing += std::memcmp(a[v]->data(), o2[6 * p + v].data(), sizeof(double) * sim_len) != 0;
    }
    std::printf("one roadsurf_run_batch: %.1f ms; runsimulation from %d threads: %.1f ms in %lld launches-batches for %lld calls; "
                "series differing: %ld\n",
                batch_ms, pool_threads, pool_ms, batches1 - batches0, calls1 - calls0, differing);
    if (differing) return 3;
  }
  int coupled = 0, failed = 0;
  for (int p = 0; p < npoints; ++p)
  {
    coupled += (status[p] & RS_ST_COUPLING_USED) != 0;
    failed += (status[p] & RS_ST_FAILED) != 0;
  }
  std::printf("coupled points %d, failed %d\n", coupled, failed);
  const PointData& d = pts[0];
  for (int i = 0; i < sim_len; i += per_hour)
    std::printf("  hour %2d  Tair %6.2f  Tsurf %7.3f  water %6.3f  ice %6.3f  snow %6.3f\n", i / per_hour, d.tair[i],
                d.Tsurf[i], d.Water[i], d.Ice[i], d.Snow[i]);
  // the surface temperature at the end of the coupling window was steered to the last observation
  const int ce = local[0].couplingIndexI;
  std::printf("Tsurf at window end %.3f, last observation %.3f\n", d.Tsurf[ce], local[0].couplingTsurf);
  return std::fabs(d.Tsurf[ce] - local[0].couplingTsurf) < 0.11 || (status[0] & RS_ST_COUPLING_FAILED) ? 0 : 2;
}
