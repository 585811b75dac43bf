// roadsurf_oracle.hpp -- CPU restatement of the RoadSurf per-point simulation loop.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under roadsurf_b200/ may include, link or call this; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the reference (fmidev/RoadSurf 1.6.1) ships no tests, golden vectors or
// expected outputs, and no Fortran compiler exists in this environment, so this restatement has
// never been compared with the gfortran build.  It follows the Fortran subroutine by
// subroutine (every function cites file:line relative to the reference tree), keeps the
// reference's AoS derived types, its per-point `do while` driver, its in-place mutation of the
// input arrays, and the REAL(4) rounding of every un-suffixed Fortran literal (macro F4).
//
// The scalar type R is a template parameter: R = double is the oracle proper, R = Counted
// (roadsurf_oracle.cpp) counts arithmetic per category for the roofline denominator.
#pragma once

#include <cmath>
#include <cstdio>
#include <vector>

#include "../include/roadsurf_b200.h"

namespace rs_oracle
{
// An un-suffixed Fortran real literal is REAL(4): it is rounded to float and then promoted.
#define F4(x) R(static_cast<double>(x##f))

// ---- math shims: overloaded for double here, for Counted in the .cpp ------------------------
inline double r_sqrt(double x) { return std::sqrt(x); }
inline double r_exp(double x) { return std::exp(x); }
inline double r_log(double x) { return std::log(x); }
inline double r_sin(double x) { return std::sin(x); }
inline double r_cos(double x) { return std::cos(x); }
inline double r_acos(double x) { return std::acos(x); }
inline double r_asin(double x) { return std::asin(x); }
inline double r_atan2(double y, double x) { return std::atan2(y, x); }
inline double r_pow(double x, double y) { return std::pow(x, y); }
inline double r_abs(double x) { return std::fabs(x); }
inline double r_aint(double x) { return std::trunc(x); }
inline double r_val(double x) { return x; }

template <class R>
inline R r_max(R a, R b)
{
  return (a > b) ? a : b;  // Fortran MAX
}
template <class R>
inline R r_min(R a, R b)
{
  return (a < b) ? a : b;  // Fortran MIN
}

// REAL(4) integer power as libgfortran/libgcc __powisf2 computes it (square and multiply).
inline float powi_f32(float x, int n)
{
  unsigned m = (n < 0) ? -static_cast<unsigned>(n) : static_cast<unsigned>(n);
  float y = (m % 2) ? x : 1.0f;
  while (m >>= 1)
  {
    x = x * x;
    if (m % 2) y = y * x;
  }
  return (n < 0) ? 1.0f / y : y;
}

constexpr int SURFACE_SNOW_DRY = 1;  // src/Constants.h
constexpr int SURFACE_SNOW_WET = 2;

// ---- derived types (src/*.f90.inc) -----------------------------------------------------------

template <class R>
struct AtmVariables  // src/AtmVariables.f90.inc
{
  R Tair, VZ, Tdew, RHz, PrecInTStep, TairInitEnd, VZInitEnd, RhzInitEnd, BLCond, RNet, LE_Flux,
      RainIntensity, SnowIntensity, TairR, VZR, RhzR, CalmLim, SensibleHeatFlux, RainmmTS,
      SnowmmTS;
  int SnowType, PrecType;
};

template <class R>
struct GroundVariables  // src/GroundVariables.f90.inc; index = Fortran index
{
  R Albedo, HStor;
  std::vector<R> condDZ, capDZ, Wcont, VSH, HS, CC, Tmp, TmpNw, DyC, DyK, ZDpth, GCond;
  R GroundFlux;
};

template <class R>
struct SurfaceVariables  // src/SurfaceVariables.f90.inc
{
  R TsurfAve, SrfWatmms, SrfSnowmms, SrfIcemms, SrfIce2mms, SrfDepmms, Q2Melt, T4Melt, TrfFric,
      EvapmmTS;
  bool VeryCold, WearSurf;
  R TsurfOBS;
};

template <class R>
struct CouplingVariables  // src/CouplingVariables.f90.inc (multi-coupling arrays cut to what
                          // the single-coupling code path touches: element 1 of each)
{
  std::vector<R> TmpSave;
  int Coupling_iterations;
  R TsurfNearestAbove, TsurfNearestBelow, RadCoeff;
  bool Down, start_coupling_again, Coupling_failed, inCouplingPhase, VeryColdSave;
  R RadCoefNearestAbove, RadCoefNearestBelow, RadCoeffPrevious, SWRadCof, LWRadCof,
      SW_correction, LW_correction, Tsurf_end_coup1;
  int couplingStartI[49], couplingEndI[49];
  R TSurfAveSave, SrfWatmmsSave, SrfIcemmsSave, SrfIce2mmsSave, SrfDepmmsSave, SrfSnowmmsSave,
      AlbedoSave, lastTsurfObs;
  int saveDatai;
  std::vector<R> SWSave, SWDirSave, LWSave;
  int NObs, CoupPhaseN;
  int obsI[49];
  R obsTsurf[49];
};

template <class R>
struct ModelSettings  // src/ModelSettings.f90.inc
{
  int InitLenI, SimLen;
  bool use_coupling, use_relaxation, force_tsurf;
  int NLayers;
  R DTSecs, tsurfOutputDepth;
  bool simulation_failed;
  R Tph, NightOn, NightOff, CalmLimDay, CalmLimNgt, TrfFricNgt, TrFfricDay;
  int coupling_minutes;
  R couplingEffectReduction;
  int outputStep;
};

template <class R>
struct PhysicalParameters  // src/PhysicalParameters.f90.inc
{
  R VK_Const, SB_const, ZRefW, ZRefT, ZeroDisp, ZMom, ZHeat, logMom, logHeat, logCond, logUstar,
      Grav, Emiss, Afc1, Bfc1, Cfc1, Dfc1, Efc1, Afc2, Bfc2, Cfc2, Dfc2, Efc2, Poro1, Poro2, vsh1,
      vsh2, LVap, LFus, TClimG, MaxPormms, DampDpth, Omega, AZ, Silt1, Silt2, RhoB1, RhoB2;
};

template <class R>
struct RoadCondParameters  // src/RoadCondParameters.f90.inc
{
  R WDampLim, WWetLim, WWearLim, Snow2IceFac, SnowIceRat, MissValI, MissValR, MinPrecmm,
      MinWatmms, MinSnowmms, MinDepmms, MinIcemms, MaxSnowmms, MaxDepmms, MaxIcemms, MaxExtmms,
      MaxWatmms, AlbDry, AlbSnow, WatDens, SnowDens, IceDens, DepDens, WatMHeat, PorEvaF,
      DampWearF, TLimFreeze, TLimMeltSnow, TLimMeltIce, TLimMeltDep, TLimDew, TLimColdH,
      TLimColdL, WetSnowFormR, WetSnowMeltR, PLimSnow, PLimRain;
  bool WetSnowFrozen;
  R freezing_limit_normal, snow_melting_limit_normal, ice_melting_limit_normal,
      frost_melting_limit_normal, frost_formation_limit_normal, T4Melt_normal;
  bool forceIceMelting, forceSnowMelting, CanMeltingChangeTemperature;
};

template <class R>
struct WearingFactors  // src/WearingFactors.f90.inc
{
  R SnowTran, DepWear, IceWear, IceWear2, WatWear;
};

// 1-based views of the caller's arrays (src/InputArrays.f90.inc, src/OutputArrays.f90.inc,
// associated by src/ConnectFortran2Carrays.f90:38-61,73-84).  Stored as double in memory; the
// accessors below convert to R on read.
struct InputArrays
{
  double *Tair, *Tdew, *VZ, *Rhz, *prec, *SW, *LW, *SW_dir, *LW_net, *TSurfObs, *local_horizons,
      *depth;
  int *PrecPhase, *year, *month, *day, *hour, *minute, *second;
};
struct OutputArrays
{
  double *TsurfOut, *SnowOut, *WaterOut, *IceOut, *DepositOut, *Ice2Out;
};

// Per-run diagnostics the reference only prints.
struct Diagnostics
{
  long long executed_steps = 0;
  long long bl_iterations = 0;
  long long bl_unstable = 0;
  int coupling_restarts = 0;
  bool bl_not_converged = false;
  bool solar_stop = false;
  bool coupling_failed = false;
  bool coupling_used = false;
  bool bad_input = false;
  bool abnormal_tsurf = false;
  bool verbose = false;
  // optional per-step trace [SimLen][TRACE_N] (last visit of a step wins); see roadModelOneStep
  double* trace = nullptr;
};
constexpr int TRACE_N = 18;

template <class R>
struct Model
{
  InputArrays modelInput;
  OutputArrays modelOutput;
  PhysicalParameters<R> phy;
  GroundVariables<R> ground;
  SurfaceVariables<R> surf;
  AtmVariables<R> atm;
  CouplingVariables<R> coupling;
  ModelSettings<R> settings;
  RoadCondParameters<R> condParam;
  Diagnostics diag;

  // ============================ src/ConnectFortran2Carrays.f90:8-86 ===========================
  void ConnectFortran2Carrays(const InputPointers& inP, OutputPointers& outP)
  {
    modelInput.Tair = inP.c_tair;
    modelInput.Tdew = inP.c_tdew;
    modelInput.VZ = inP.c_VZ;
    modelInput.Rhz = inP.c_Rhz;
    modelInput.prec = inP.c_prec;
    modelInput.SW = inP.c_SW;
    modelInput.LW = inP.c_LW;
    modelInput.SW_dir = inP.c_SW_dir;
    modelInput.LW_net = inP.c_LW_net;
    modelInput.TSurfObs = inP.c_TSurfObs;
    modelInput.PrecPhase = inP.c_PrecPhase;
    modelInput.local_horizons = inP.c_local_horizons;
    modelInput.depth = inP.c_Depth;
    modelInput.year = inP.c_year;
    modelInput.month = inP.c_month;
    modelInput.day = inP.c_day;
    modelInput.hour = inP.c_hour;
    modelInput.minute = inP.c_minute;
    modelInput.second = inP.c_second;
    modelOutput.TsurfOut = outP.c_TsurfOut;
    modelOutput.SnowOut = outP.c_SnowOut;
    modelOutput.WaterOut = outP.c_WaterOut;
    modelOutput.IceOut = outP.c_IceOut;
    modelOutput.DepositOut = outP.c_DepositOut;
    modelOutput.Ice2Out = outP.c_Ice2Out;
  }

  // =================================== src/Initialization.f90 =================================

  static bool int2Logical(int v) { return v == 1; }  // :560-571

  // :442-476
  void initSettings(const InputSettings& inS, const InputParameters& ip, const LocalParameters& lp)
  {
    settings.SimLen = inS.SimLen;
    settings.InitLenI = lp.InitLenI;
    settings.DTSecs = R(inS.DTSecs);
    settings.tsurfOutputDepth = R(inS.tsurfOutputDepth);
    settings.NLayers = inS.NLayers;
    settings.NightOn = R(ip.NightOn);
    settings.NightOff = R(ip.NightOff);
    settings.CalmLimDay = R(ip.CalmLimDay);
    settings.CalmLimNgt = R(ip.CalmLimNgt);
    settings.TrfFricNgt = R(ip.TrfFricNgt);
    settings.TrFfricDay = R(ip.TrFfricDay);
    settings.use_coupling = int2Logical(inS.use_coupling);
    settings.use_relaxation = int2Logical(inS.use_relaxation);
    settings.coupling_minutes = inS.coupling_minutes;
    settings.couplingEffectReduction = R(inS.couplingEffectReduction);
    settings.outputStep = inS.outputStep;
    settings.force_tsurf = int2Logical(inS.force_tsurf);
  }

  // :397-412
  void initOutputArrays(int SimLen)
  {
    for (int i = 1; i <= SimLen; ++i)
    {
      modelOutput.SnowOut[i - 1] = -9999.0;
      modelOutput.WaterOut[i - 1] = -9999.0;
      modelOutput.IceOut[i - 1] = -9999.0;
      modelOutput.Ice2Out[i - 1] = -9999.0;
      modelOutput.DepositOut[i - 1] = -9999.0;
      modelOutput.TsurfOut[i - 1] = -9999.0;
    }
  }

  // src/InputOutput.f90:4-39 (+ initTsurfObsArrays, src/Initialization.f90:417-439)
  void setInputParam(const LocalParameters& lp)
  {
    coupling.NObs = -99;
    atm.TairR = R(static_cast<double>(static_cast<float>(lp.tair_relax)));
    atm.VZR = R(static_cast<double>(static_cast<float>(lp.VZ_relax)));
    atm.RhzR = R(static_cast<double>(static_cast<float>(lp.RH_relax)));
    if (atm.TairR < F4(-100.0) || atm.TairR > F4(100.0) || atm.VZR < F4(0.0) ||
        atm.VZR > F4(100.0) || atm.RhzR < F4(0.0) || atm.RhzR > R(110))
    {
      settings.use_relaxation = false;
    }
    for (int i = 1; i <= 48; ++i)
    {
      coupling.obsI[i] = -99;
      coupling.obsTsurf[i] = F4(-99.0);
    }
    coupling.obsI[1] = lp.couplingIndexI;
    coupling.obsTsurf[1] = R(lp.couplingTsurf);
    coupling.lastTsurfObs = R(lp.couplingTsurf);
    coupling.NObs = 1;
    if (R(lp.couplingTsurf) < R(-100) || coupling.obsI[1] < 1)
    {
      settings.use_coupling = false;
    }
  }

  // :150-178
  void allocator()
  {
    const int n = settings.NLayers;
    ground.condDZ.assign(n + 2, R(0));
    ground.capDZ.assign(n + 2, R(0));
    ground.Wcont.assign(n + 2, R(0));
    ground.VSH.assign(n + 2, R(0));
    ground.HS.assign(n + 2, R(0));
    ground.CC.assign(n + 2, R(0));
    ground.Tmp.assign(n + 2, R(0));
    ground.TmpNw.assign(n + 2, R(0));
    ground.DyC.assign(n + 2, R(0));
    ground.DyK.assign(n + 2, R(0));
    ground.ZDpth.assign(n + 2, R(0));
    ground.GCond.assign(n + 2, R(0));
    coupling.TmpSave.assign(n + 2, R(0));
  }

  // :217-235.  0.0103*1.4**(I-1) is a REAL(4) expression; ZAdd = 0.02 is REAL(4) stored in REAL(8).
  void initDepth()
  {
    const int n = settings.NLayers;
    const R ZAdd = F4(0.02);
    ground.ZDpth[1] = R(0.0);
    for (int I = 1; I <= n; ++I)
    {
      const float inc = 0.0103f * powi_f32(1.4f, I - 1);
      ground.ZDpth[I + 1] = ground.ZDpth[I] + R(static_cast<double>(inc)) + ZAdd;
    }
  }

  // :290-308
  void initSurf(bool wearOn)
  {
    surf.Q2Melt = R(0.0);
    surf.VeryCold = false;
    surf.WearSurf = wearOn;
    surf.TrfFric = R(5.0);
    surf.EvapmmTS = R(0.0);
    surf.TsurfOBS = F4(-99.9);
    surf.SrfWatmms = R(0.0);
    surf.SrfSnowmms = R(0.0);
    surf.SrfIcemms = R(0.0);
    surf.SrfIce2mms = R(0.0);
    surf.SrfDepmms = R(0.0);
  }

  // :310-358
  void InitParam(const InputParameters& ip)
  {
    phy.Grav = R(ip.Grav);
    phy.SB_const = R(ip.SB_Const);
    phy.VK_Const = R(ip.VK_Const);
    phy.ZRefW = R(ip.ZRefW);
    phy.ZRefT = R(ip.ZRefT);
    phy.ZeroDisp = R(ip.ZeroDisp);
    phy.ZMom = R(ip.ZMom);
    phy.ZHeat = R(ip.ZHeat);
    phy.logMom = r_log((phy.ZRefW + phy.ZMom) / phy.ZMom);
    phy.logHeat = r_log((phy.ZRefW + phy.ZHeat) / phy.ZHeat);
    phy.logCond = r_log((phy.ZRefW - phy.ZeroDisp + phy.ZHeat) / phy.ZHeat);
    phy.logUstar = r_log((phy.ZRefW - phy.ZeroDisp + phy.ZMom) / phy.ZMom);
    phy.Emiss = R(ip.Emiss);
    ground.Albedo = R(ip.Albedo);
    phy.MaxPormms = R(ip.MaxPormms);
    phy.TClimG = R(ip.TClimG);
    phy.DampDpth = R(ip.DampDpth);
    phy.Omega = R(ip.Omega);
    phy.AZ = R(ip.AZ);
    phy.LVap = R(ip.LVap);
    phy.LFus = R(ip.LFus);
    phy.vsh1 = R(ip.vsh1);
    phy.vsh2 = R(ip.vsh2);
    phy.Poro1 = R(ip.Poro1);
    phy.Poro2 = R(ip.Poro2);
    phy.RhoB1 = R(ip.RhoB1);
    phy.RhoB2 = R(ip.RhoB2);
    phy.Silt1 = R(ip.Silt1);
    phy.Silt2 = R(ip.Silt2);
  }

  // src/BalanceModel.f90:325-351
  int JulDay(int i) const
  {
    static const int MonEnd[25] = {0,  0,  31, 59, 90,  120, 151, 181, 212, 243, 273, 304, 334,
                                   0,  31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335};
    const int syear = modelInput.year[i - 1];
    const int smon = modelInput.month[i - 1];
    const int sday = modelInput.day[i - 1];
    auto imin = [](int a, int b) { return a < b ? a : b; };
    const int leapcorr =
        1 - imin(syear % 4, 1) + imin(syear % 100, 1) - imin(syear % 400, 1);
    return MonEnd[smon + leapcorr * 12] + sday;
  }

  // src/BalanceModel.f90:390-417
  R getTempAtDepth(R depth) const
  {
    const int zlen = settings.NLayers + 1;  // size(ground%ZDpth)
    if (r_abs(depth - F4(0.0)) < F4(0.00001))
    {
      return ground.Tmp[1];
    }
    else if (depth > ground.ZDpth[zlen])
    {
      return ground.Tmp[zlen];
    }
    int idx;
    for (idx = 1; idx <= zlen - 1; ++idx)
    {
      if (depth > ground.ZDpth[idx] && depth <= ground.ZDpth[idx + 1]) break;
    }
    if (idx > zlen - 1) idx = zlen - 1;  // unreachable for depth >= 1e-5; keeps the read in bounds
    return ground.Tmp[idx] + (depth - ground.ZDpth[idx]) * (ground.Tmp[idx + 1] - ground.Tmp[idx]) /
                                 (ground.ZDpth[idx + 1] - ground.ZDpth[idx]);
  }

  // :238-287
  void initTemp(R Tsurf, R Tair, R depth)
  {
    const int n = settings.NLayers;
    ground.Tmp[0] = Tair;
    if (Tsurf > R(-100))
    {
      for (int i = 1; i <= 4; ++i) ground.Tmp[i] = Tsurf;
    }
    else
    {
      for (int i = 1; i <= 4; ++i) ground.Tmp[i] = Tair;
    }
    const int juld = JulDay(1);
    ground.Tmp[n + 1] =
        phy.TClimG + phy.AZ * r_sin(phy.Omega * R(juld) + phy.Omega * R(-170) -
                                    (ground.ZDpth[n + 1] / phy.DampDpth));
    for (int i = 5; i <= n; ++i)
    {
      ground.Tmp[i] = ground.Tmp[4] + (ground.Tmp[n + 1] - ground.Tmp[4]) /
                                          (ground.ZDpth[n + 1] - ground.ZDpth[4]) *
                                          (ground.ZDpth[i] - ground.ZDpth[4]);
    }
    for (int i = 0; i <= n + 1; ++i) ground.TmpNw[i] = ground.Tmp[i];
    if (depth >= R(0))
    {
      surf.TsurfAve = getTempAtDepth(depth);
    }
    else
    {
      surf.TsurfAve = R(0.5) * (ground.Tmp[1] + ground.Tmp[2]);
    }
  }

  // :360-394
  void initVariables()
  {
    atm.Tair = F4(-99.9);
    atm.VZ = F4(-99.9);
    atm.RHz = F4(-99.9);
    atm.PrecInTStep = F4(-99.9);
    atm.BLCond = F4(-99.9);
    atm.TairInitEnd = F4(-99.9);
    atm.VZInitEnd = F4(-99.9);
    atm.RhzInitEnd = F4(-99.9);
    atm.RainmmTS = R(0.0);
    atm.SnowmmTS = R(0.0);
    atm.SnowType = SURFACE_SNOW_DRY;
    atm.CalmLim = F4(0.4);
    atm.SensibleHeatFlux = F4(-9999.9);
    ground.GroundFlux = F4(-9999.9);
    for (int I = 1; I <= settings.NLayers; ++I)
    {
      ground.VSH[I] = F4(-99.9);
      ground.HS[I] = F4(-99.9);
      ground.CC[I] = F4(-99.9);
      ground.GCond[I] = F4(-99.9);
    }
    ground.GCond[0] = F4(-99.9);
  }

  // src/BalanceModel.f90:158-186
  void HCapValues()
  {
    phy.Afc1 = F4(0.65) - F4(0.78) * phy.RhoB1 + F4(0.60) * phy.RhoB1 * phy.RhoB1;
    phy.Bfc1 = F4(1.06) * phy.RhoB1;
    if (phy.Silt1 > F4(0.00001))
      phy.Cfc1 = R(1) + F4(2.6) / r_sqrt(phy.Silt1);
    else
      phy.Cfc1 = R(0.);
    phy.Dfc1 = F4(0.03) + F4(0.1) * phy.RhoB1 * phy.RhoB1;
    phy.Efc1 = R(4);
    phy.Afc2 = F4(0.65) - F4(0.78) * phy.RhoB2 + F4(0.60) * phy.RhoB2 * phy.RhoB2;
    phy.Bfc2 = F4(1.06) * phy.RhoB2;
    if (phy.Silt2 > F4(0.00001))
      phy.Cfc2 = R(1) + F4(2.6) / r_sqrt(phy.Silt2);
    else
      phy.Cfc2 = R(0.);
    phy.Dfc2 = F4(0.03) + F4(0.1) * phy.RhoB2 * phy.RhoB2;
    phy.Efc2 = R(4);
  }

  // :181-214
  void ground_prop_init()
  {
    const int n = settings.NLayers;
    ground.DyC[1] = (ground.ZDpth[2] - ground.ZDpth[1]) / R(2.0);
    for (int j = 2; j <= n; ++j) ground.DyC[j] = (ground.ZDpth[j + 1] - ground.ZDpth[j - 1]) / R(2.0);
    for (int j = 1; j <= n; ++j) ground.DyK[j] = ground.ZDpth[j + 1] - ground.ZDpth[j];
    ground.Wcont[1] = F4(0.01);
    ground.Wcont[2] = F4(0.01);
    ground.Wcont[3] = F4(0.3);
    ground.Wcont[4] = F4(0.3);
    for (int I = 5; I <= n; ++I) ground.Wcont[I] = F4(0.3);
  }

  // src/BalanceModel.f90:254-279
  void CalcCC()
  {
    for (int I = 1; I <= settings.NLayers; ++I)
    {
      if (I <= 2)
      {
        ground.CC[I] = phy.Afc1 + phy.Bfc1 * ground.Wcont[I] -
                       (phy.Afc1 - phy.Dfc1) * r_exp(-r_pow(phy.Cfc1 * ground.Wcont[I], phy.Efc1));
      }
      else
      {
        ground.CC[I] = phy.Afc2 + phy.Bfc2 * ground.Wcont[I] -
                       (phy.Afc2 - phy.Dfc2) * r_exp(-r_pow(phy.Cfc2 * ground.Wcont[I], phy.Efc2));
      }
    }
  }

  // src/BalanceModel.f90:189-251
  void CalcHCapHCond()
  {
    const R DTSecs = settings.DTSecs;
    ground.GCond[0] = atm.BLCond;
    for (int I = 1; I <= settings.NLayers; ++I)
    {
      R RooWT, CWT;
      if (ground.TmpNw[I] >= R(0))
      {
        const R tmp2 = ground.TmpNw[I] * ground.TmpNw[I];
        RooWT = -F4(0.0050) * tmp2 + F4(0.0079) * ground.TmpNw[I] + F4(1000.0028);
        CWT = F4(0.0000102) * tmp2 * tmp2 - F4(0.0017169) * tmp2 * ground.TmpNw[I] +
              F4(0.11516) * tmp2 - F4(3.4739) * ground.TmpNw[I] + F4(4217.2);
      }
      else
      {
        RooWT = F4(920.0);
        CWT = F4(2100.0);
      }
      const R CHWT = RooWT * CWT;
      if (I <= 2)
        ground.VSH[I] = (R(1.0) - phy.Poro1) * phy.vsh1 + ground.Wcont[I] * CHWT;
      else
        ground.VSH[I] = (R(1.0) - phy.Poro2) * phy.vsh2 + ground.Wcont[I] * CHWT;
      if (I == 1)
        ground.HS[I] = ground.VSH[I] * (ground.ZDpth[I + 1] - ground.ZDpth[I]) / (R(2.0) * DTSecs);
      else
        ground.HS[I] =
            ground.VSH[I] * (ground.ZDpth[I + 1] - ground.ZDpth[I - 1]) / (R(2.0) * DTSecs);
      ground.GCond[I] = ground.CC[I] / (ground.ZDpth[I + 1] - ground.ZDpth[I]);
    }
  }

  // src/BalanceModel.f90:132-155
  void calcCapDZCondDZ()
  {
    for (int j = 1; j <= settings.NLayers; ++j)
    {
      ground.condDZ[j] = -(ground.CC[j] / ground.DyK[j]);
      ground.capDZ[j] = -(R(1) / (ground.DyC[j] * ground.VSH[j]));
    }
  }

  // src/Coupling.f90:144-169
  void initCoupling()
  {
    coupling.Coupling_iterations = 0;
    coupling.TsurfNearestAbove = R(-9999.0);
    coupling.TsurfNearestBelow = R(-9999.0);
    coupling.RadCoeff = R(1.0);
    coupling.RadCoefNearestAbove = R(-9999.0);
    coupling.RadCoefNearestBelow = R(-9999.0);
    coupling.RadCoeffPrevious = R(1.0);
    coupling.SWRadCof = R(1.0);
    coupling.LWRadCof = R(1.0);
    coupling.start_coupling_again = false;
    coupling.Coupling_failed = false;
    coupling.SW_correction = R(0.0);
    coupling.LW_correction = R(0.0);
    coupling.inCouplingPhase = false;
    coupling.CoupPhaseN = 1;
    coupling.lastTsurfObs = coupling.obsTsurf[1];
  }

  // src/Coupling.f90:486-534
  void initCouplingTimes()
  {
    const R DTs = settings.DTSecs;
    for (int i = 1; i <= 48; ++i)
    {
      coupling.couplingStartI[i] = -99;
      coupling.couplingEndI[i] = -99;
    }
    if (settings.use_coupling && coupling.obsI[1] > -1)
    {
      coupling.couplingEndI[1] = coupling.obsI[1];
      if (R(coupling.obsI[1]) <= R(settings.coupling_minutes * 60) / DTs)
      {
        coupling.couplingStartI[1] = 1;
      }
      else
      {
        coupling.couplingStartI[1] =
            coupling.obsI[1] - static_cast<int>(r_val(R(settings.coupling_minutes * 60) / DTs));
      }
    }
    else
    {
      settings.use_coupling = false;
    }
    int couplingLen = coupling.couplingEndI[1] - coupling.couplingStartI[1] + 1;
    if (couplingLen < 0) couplingLen = 0;
    coupling.SWSave.assign(couplingLen + 1, F4(-9999.9));
    coupling.SWDirSave.assign(couplingLen + 1, F4(-9999.9));
    coupling.LWSave.assign(couplingLen + 1, F4(-9999.9));
  }

  // :479-557
  void condInit(const InputParameters& ip)
  {
    RoadCondParameters<R>& c = condParam;
    c.WatDens = R(ip.WatDens);
    c.SnowDens = R(ip.SnowDens);
    c.IceDens = R(ip.IceDens);
    c.DepDens = R(ip.DepDens);
    c.WatMHeat = R(ip.WatMHeat);
    c.PorEvaF = R(ip.PorEvaF);
    c.DampWearF = R(ip.DampWearF);
    c.freezing_limit_normal = R(ip.freezing_limit_normal);
    c.snow_melting_limit_normal = R(ip.snow_melting_limit_normal);
    c.ice_melting_limit_normal = R(ip.ice_melting_limit_normal);
    c.frost_melting_limit_normal = R(ip.frost_melting_limit_normal);
    c.frost_formation_limit_normal = R(ip.frost_formation_limit_normal);
    c.T4Melt_normal = R(ip.T4Melt_normal);
    c.TLimFreeze = R(ip.freezing_limit_normal);
    c.TLimMeltSnow = R(ip.snow_melting_limit_normal);
    c.TLimMeltIce = R(ip.ice_melting_limit_normal);
    c.TLimMeltDep = R(ip.frost_melting_limit_normal);
    c.TLimDew = R(ip.frost_formation_limit_normal);
    surf.T4Melt = R(ip.T4Melt_normal);
    c.TLimColdH = R(ip.TLimColdH);
    c.TLimColdL = R(ip.TLimColdL);
    c.WetSnowFormR = R(ip.WetSnowFormR);
    c.WetSnowMeltR = R(ip.WetSnowMeltR);
    c.PLimSnow = R(ip.PLimSnow);
    c.PLimRain = R(ip.PLimRain);
    c.MinPrecmm = R(ip.MinPrecmm);
    c.MinWatmms = R(ip.MinWatmms);
    c.MinSnowmms = R(ip.MinSnowmms);
    c.MinDepmms = R(ip.MinDepmms);
    c.MinIcemms = R(ip.MinIcemms);
    c.MaxSnowmms = R(ip.MaxSnowmms);
    c.MaxDepmms = R(ip.MaxDepmms);
    c.MaxIcemms = R(ip.MaxIcemms);
    c.MaxExtmms = R(ip.MaxExtmms);
    c.MaxWatmms = R(ip.MaxWatmms);
    c.AlbDry = R(ip.AlbDry);
    c.AlbSnow = R(ip.AlbSnow);
    c.MissValI = R(ip.MissValI);
    c.MissValR = R(ip.MissValR);
    c.WDampLim = R(ip.WDampLim);
    c.WWetLim = R(ip.WWetLim);
    c.WWearLim = R(ip.WWearLim);
    c.Snow2IceFac = R(ip.Snow2IceFac);
    c.SnowIceRat = R(0.0);
    c.WetSnowFrozen = false;
    c.forceIceMelting = false;
    c.forceSnowMelting = false;
    c.CanMeltingChangeTemperature = true;
  }

  // :66-147
  void initVariablesAndParameters(const InputParameters& ip)
  {
    settings.simulation_failed = false;
    settings.Tph = settings.DTSecs / R(3600.0);
    allocator();
    initDepth();
    initSurf(true);
    InitParam(ip);
    initTemp(R(modelInput.TSurfObs[0]), R(modelInput.Tair[0]), R(modelInput.depth[0]));
    initVariables();
    HCapValues();
    ground_prop_init();
    CalcCC();
    CalcHCapHCond();
    calcCapDZCondDZ();
    initCoupling();
    initCouplingTimes();
    condInit(ip);
    if (R(modelInput.VZ[0]) < F4(0.4))
    {
      modelInput.VZ[0] = static_cast<double>(0.4f);
    }
    atm.Tair = R(modelInput.Tair[0]);
    atm.VZ = R(modelInput.VZ[0]);
    atm.RHz = R(modelInput.Rhz[0]);
    const R depth = R(modelInput.depth[0]);
    if (depth >= R(0))
      surf.TsurfAve = getTempAtDepth(depth);
    else
      surf.TsurfAve = (ground.Tmp[1] + ground.Tmp[2]) / R(2.0);
    CalcBLCondAndLE();
    if (coupling.lastTsurfObs < R(-100)) coupling.Coupling_failed = true;
  }

  // :9-62
  void Initialization(const InputSettings& inS, const InputParameters& ip, const LocalParameters& lp)
  {
    initSettings(inS, ip, lp);
    initOutputArrays(settings.SimLen);
    setInputParam(lp);
    initVariablesAndParameters(ip);
  }

  // =================================== src/BoundaryLayer.f90 ==================================

  // :112-131
  static R calcRaero(R logMom, R logHeat, R PSIM, R PSIH, R VK_Const, R VZ)
  {
    R RAero = (logMom + PSIM) * (logHeat + PSIH) / (VK_Const * VK_Const * VZ);
    if (RAero > R(30.0)) RAero = R(30.);
    return RAero;
  }

  // :134-190
  static void CalcLE(R TSurfAve, R TAmb, R Rhz, R AirDens, R AirHCap, R PsychC, R RAero, R LVap,
                     R LFus, R WatDen, R DtSecs, R SrfWatmms, R& LE_Flux, R& EvapmmTS)
  {
    R ESat;
    if (TSurfAve < R(0))
      ESat = F4(0.61078) * r_exp(F4(21.875) * TSurfAve / (TSurfAve + F4(265.5)));
    else
      ESat = F4(0.61078) * r_exp(F4(17.269) * TSurfAve / (TSurfAve + F4(237.3)));
    const R ESurf = ESat;
    if (TAmb < R(0))
      ESat = F4(0.61078) * r_exp(F4(21.875) * TAmb / (TAmb + F4(265.5)));
    else
      ESat = F4(0.61078) * r_exp(F4(17.269) * TAmb / (TAmb + F4(237.3)));
    const R EAir = r_min(F4(0.01) * Rhz, R(1.0)) * ESat;
    LE_Flux = (AirDens * AirHCap * (ESurf - EAir)) / (PsychC * RAero);
    if (TSurfAve >= R(0.0))
      EvapmmTS = (LE_Flux / (LVap * WatDen)) * R(1000.0) * DtSecs;
    else
      EvapmmTS = (LE_Flux / (LFus * WatDen)) * R(1000.0) * DtSecs;
    if (LE_Flux > R(0.0) && SrfWatmms <= R(0.0))
    {
      LE_Flux = R(0.0);
      EvapmmTS = R(0.0);
    }
  }

  // :3-109
  void CalcBLCondAndLE()
  {
    const R TSurfAve = surf.TsurfAve;
    const R ConvLim = F4(0.001);
    const int MaxIter = 40;
    const R Tair = atm.Tair;
    const R VZ = atm.VZ;
    const R Rhz = atm.RHz;
    R BLCond = atm.BLCond;
    const R TaK = Tair + F4(273.15);
    const R AirDens = R(100000.0) / (F4(287.05) * TaK);
    const R AirHCap = R(1005.0) + ((TaK - R(250.0)) * (TaK - R(250.0))) / R(3364.);
    const R AirVCap = AirHCap * AirDens;
    const R PsychC = F4(0.1) * (F4(0.00063) * TaK + F4(0.47496));
    const R WatDen = -F4(0.0050) * TSurfAve * TSurfAve + F4(0.0079) * TSurfAve + F4(1000.0028);
    R PSIM = R(0.0), PSIH = R(0.0), Stab = R(0.0), UStar, BLCond_Old = BLCond;
    int j;
    for (j = 1; j <= MaxIter; ++j)
    {
      BLCond_Old = BLCond;
      UStar = phy.VK_Const * VZ / (phy.logUstar + PSIM);
      BLCond = AirVCap * phy.VK_Const * UStar / (phy.logCond + PSIH);
      Stab = -phy.VK_Const * phy.ZRefT * phy.Grav * BLCond * (TSurfAve - Tair) /
             (AirVCap * (Tair + F4(273.15)) * (UStar * UStar * UStar));
      if (Stab > R(1)) Stab = R(1);
      if (Stab > R(0))
      {
        PSIH = F4(4.7) * Stab;
        PSIM = PSIH;
      }
      else
      {
        PSIH = -R(2.0) * r_log((R(1.0) + r_sqrt(R(1.0) - R(16.0) * Stab)) / R(2.0));
        PSIM = F4(0.6) * PSIH;
        ++diag.bl_unstable;
      }
      ++diag.bl_iterations;
      if (r_abs(BLCond - BLCond_Old) < ConvLim && j >= 5) break;
    }
    // Fortran leaves j = MaxIter+1 after a loop that ran to completion.
    if (r_abs(BLCond - BLCond_Old) > R(10) * ConvLim && j >= 5) diag.bl_not_converged = true;
    const R RAero = calcRaero(phy.logMom, phy.logHeat, PSIM, PSIH, phy.VK_Const, VZ);
    CalcLE(TSurfAve, Tair, Rhz, AirDens, AirHCap, PsychC, RAero, phy.LVap, phy.LFus, WatDen,
           settings.DTSecs, surf.SrfWatmms, atm.LE_Flux, surf.EvapmmTS);
    atm.BLCond = BLCond;
  }

  // =================================== src/InputOutput.f90 ====================================

  bool sky_view_active(const LocalParameters& lp) const
  {
    return R(lp.sky_view) < R(1.0) && R(lp.sky_view) > F4(-0.01);
  }

  // :45-84
  void CheckValues(int i, const LocalParameters& lp)
  {
    const int k = i - 1;
    const InputArrays& m = modelInput;
    if (R(m.Tair[k]) < R(-90.0) || R(m.Tair[k]) > R(100.0) || R(m.Tdew[k]) < R(-90) ||
        R(m.Tdew[k]) > R(100.0) || R(m.Rhz[k]) < F4(-0.1) || R(m.Rhz[k]) > R(120.0) ||
        R(m.VZ[k]) < R(-1.0) || R(m.VZ[k]) > R(100.0) || R(m.SW[k]) < F4(-0.1) ||
        R(m.SW[k]) > R(4000.0) || R(m.LW[k]) < F4(-0.1) || R(m.LW[k]) > R(1000.0) ||
        R(m.prec[k]) < F4(-0.1) || R(m.prec[k]) > R(500.0))
    {
      if (diag.verbose) std::printf(" BAD input value! step %d\n", i);
      diag.bad_input = true;
      settings.simulation_failed = true;
    }
    if (sky_view_active(lp))
    {
      if (R(m.SW_dir[k]) < F4(-0.1) || R(m.SW_dir[k]) > R(4000.0) || R(m.LW_net[k]) < R(-1000.0) ||
          R(m.LW_net[k]) > R(1000.0))
      {
        if (diag.verbose) std::printf(" BAD input value: SW_dir,LW_net step %d\n", i);
        diag.bad_input = true;
        settings.simulation_failed = true;
      }
    }
    if (m.SW_dir[k] > m.SW[k]) m.SW_dir[k] = m.SW[k];
    if (surf.TsurfAve < R(-100.0) || surf.TsurfAve > R(100.0))
    {
      if (diag.verbose) std::printf(" Abnormal surface temperature step %d\n", i);
      diag.abnormal_tsurf = true;
      settings.simulation_failed = true;
    }
  }

  // :86-149
  void SetCurrentValues(int i)
  {
    const int k = i - 1;
    atm.Tair = R(modelInput.Tair[k]);
    atm.Tdew = R(modelInput.Tdew[k]);
    atm.VZ = R(modelInput.VZ[k]);
    atm.RHz = R(modelInput.Rhz[k]);
    atm.PrecInTStep = R(modelInput.prec[k]) / R(3600) * settings.DTSecs;
    ground.Tmp[0] = atm.Tair;
    if (i <= settings.InitLenI || settings.force_tsurf)
    {
      if (R(modelInput.TSurfObs[k]) > R(-100.0))
      {
        if (!settings.use_coupling || i < coupling.couplingStartI[coupling.CoupPhaseN])
        {
          surf.TsurfOBS = R(modelInput.TSurfObs[k]);
          ground.Tmp[1] = surf.TsurfOBS;
          ground.Tmp[2] = surf.TsurfOBS;
          R depth;
          if (settings.tsurfOutputDepth >= R(0.0))
            depth = settings.tsurfOutputDepth;
          else
            depth = R(modelInput.depth[k]);
          if (depth >= R(0))
            surf.TsurfAve = getTempAtDepth(depth);
          else
            surf.TsurfAve = (ground.Tmp[1] + ground.Tmp[2]) / R(2.0);
        }
        else
        {
          surf.TsurfOBS = R(-9999.0);
        }
      }
      else
      {
        surf.TsurfOBS = R(-9999.0);
      }
    }
  }

  // :151-165
  void SaveOutput(int i)
  {
    modelOutput.SnowOut[i - 1] = r_val(surf.SrfSnowmms);
    modelOutput.WaterOut[i - 1] = r_val(surf.SrfWatmms);
    modelOutput.IceOut[i - 1] = r_val(surf.SrfIcemms);
    modelOutput.Ice2Out[i - 1] = r_val(surf.SrfIce2mms);
    modelOutput.DepositOut[i - 1] = r_val(surf.SrfDepmms);
    modelOutput.TsurfOut[i - 1] = r_val(surf.TsurfAve);
  }

  // :169-198
  void lastValues()
  {
    const int k = settings.SimLen - 1;
    atm.Tair = R(modelInput.Tair[k]);
    atm.Tdew = R(modelInput.Tdew[k]);
    atm.VZ = R(modelInput.VZ[k]);
    atm.RHz = R(modelInput.Rhz[k]);
    atm.PrecInTStep = R(modelInput.prec[k]) / R(3600) * settings.DTSecs;
    ground.Tmp[0] = atm.Tair;
    const R depth = R(modelInput.depth[k]);
    if (depth >= R(0))
      surf.TsurfAve = getTempAtDepth(depth);
    else
      surf.TsurfAve = (ground.Tmp[1] + ground.Tmp[2]) / R(2.0);
  }

  // :239-268
  static R CalcTDew(R T2m, R Rhz)
  {
    const R AFact = F4(0.61078), Alphai = F4(21.875), Betai = F4(265.5), Alphaw = F4(17.269),
            Betaw = F4(237.3);
    R Alpha, Beta;
    if (T2m >= R(0.))
    {
      Alpha = Alphaw;
      Beta = Betaw;
    }
    else
    {
      Alpha = Alphai;
      Beta = Betai;
    }
    const R EPrSat = AFact * r_exp(Alpha * T2m / (T2m + Beta));
    const R Epr = F4(0.01) * Rhz * EPrSat;
    const R XX = r_log(Epr / AFact);
    return Beta * XX / (Alpha - XX);
  }

  // :202-236 (unused by the model; kept for the Tdew<->RH identity test)
  static R CalcRhOne(R T2m, R Tdew)
  {
    const R AFact = F4(0.61078), Alphai = F4(21.875), Betai = F4(265.5), Alphaw = F4(17.269),
            Betaw = F4(237.3);
    R Alpha, Beta;
    if (T2m >= R(0.))
    {
      Alpha = Alphaw;
      Beta = Betaw;
    }
    else
    {
      Alpha = Alphai;
      Beta = Betai;
    }
    const R ESatT = AFact * r_exp(Alpha * T2m / (T2m + Beta));
    const R ESatTD = AFact * r_exp(Alpha * Tdew / (Tdew + Beta));
    return r_min((ESatTD / ESatT) * R(100.0), R(100.0));
  }

  // =================================== src/Relaxation.f90:10-47 ===============================
  void RelaxationOperations(int i)
  {
    const R DTs = settings.DTSecs;
    const int initLI = settings.InitLenI;
    if (i == initLI)
    {
      atm.TairInitEnd = atm.Tair;
      atm.VZInitEnd = atm.VZ;
      atm.RhzInitEnd = atm.RHz;
    }
    if (i > initLI)
    {
      const float four_hours = 4.f * 3600.f;  // REAL(4) literal expression
      atm.Tair = atm.Tair - (atm.TairR - atm.TairInitEnd) *
                                r_exp(-((DTs * R(i)) - (DTs * R(initLI))) / R(static_cast<double>(four_hours)));
      ground.Tmp[0] = atm.Tair;
      atm.VZ = atm.VZ - (atm.VZR - atm.VZInitEnd) *
                            r_exp(-((DTs * R(i)) - (DTs * R(initLI))) / R(static_cast<double>(four_hours)));
      atm.RHz = atm.RHz - (atm.RhzR - atm.RhzInitEnd) *
                              r_exp(-((DTs * R(i)) - (DTs * R(initLI))) / R(static_cast<double>(four_hours)));
      if (atm.RHz > R(100.)) atm.RHz = R(100.0);
    }
    atm.Tdew = CalcTDew(atm.Tair, atm.RHz);
  }

  // =================================== src/Cond.f90 ===========================================

  // :143-249
  void CalcPrecType(int PrecPhase)
  {
    const RoadCondParameters<R>& CP = condParam;
    atm.RainmmTS = R(0.0);
    atm.SnowmmTS = R(0.0);
    atm.PrecType = -1;
    bool UseInterpr = true;
    if (R(PrecPhase) > CP.MissValI)
    {
      UseInterpr = false;
      if (atm.PrecInTStep <= CP.MinPrecmm)
      {
        atm.PrecInTStep = R(0.0);
        atm.PrecType = -1;
        atm.RainmmTS = R(0.0);
        atm.SnowmmTS = R(0.0);
      }
      else
      {
        switch (PrecPhase)
        {
          case 0:
          case 1:
          case 4:
          case 5:
            atm.RainmmTS = atm.PrecInTStep;
            atm.SnowmmTS = R(0.0);
            atm.PrecType = 1;
            atm.SnowType = SURFACE_SNOW_WET;
            break;
          case 2:
            atm.SnowmmTS = atm.PrecInTStep / R(2.);
            atm.RainmmTS = atm.SnowmmTS;
            atm.PrecType = 2;
            atm.SnowType = SURFACE_SNOW_WET;
            break;
          case 3:
          case 6:
            atm.SnowmmTS = atm.PrecInTStep;
            atm.PrecType = 3;
            atm.RainmmTS = R(0.0);
            break;
          default:
            UseInterpr = true;
        }
      }
    }
    if (UseInterpr)
    {
      if (atm.PrecInTStep <= CP.MinPrecmm)
      {
        atm.PrecInTStep = R(0.0);
        atm.PrecType = -1;
        atm.RainmmTS = R(0.0);
        atm.SnowmmTS = R(0.0);
      }
      else
      {
        atm.SnowmmTS = R(0.0);
        const R PExp = R(22.0) - F4(2.7) * atm.Tair - F4(0.20) * atm.RHz;
        const R PRain = R(1.0) / (R(1.0) + r_exp(PExp));
        if (PRain < CP.PLimSnow)
        {
          atm.SnowmmTS = atm.PrecInTStep;
          atm.PrecType = 3;
        }
        else if (PRain > CP.PLimRain)
        {
          atm.RainmmTS = atm.PrecInTStep;
          atm.SnowType = SURFACE_SNOW_WET;
          atm.PrecType = 1;
        }
        else
        {
          atm.SnowmmTS = atm.PrecInTStep / R(2.);
          atm.RainmmTS = atm.SnowmmTS;
          atm.SnowType = SURFACE_SNOW_WET;
          atm.PrecType = 2;
        }
      }
    }
    atm.RainIntensity = (atm.RainmmTS / settings.DTSecs) * R(3600.0);
    atm.SnowIntensity = (atm.SnowmmTS / settings.DTSecs) * R(3600.0);
  }

  // src/Storage.f90:9-29
  void PrecipitationToStorage(int PrecPhase)
  {
    CalcPrecType(PrecPhase);
    surf.SrfWatmms = surf.SrfWatmms + atm.RainmmTS;
    surf.SrfSnowmms = surf.SrfSnowmms + atm.SnowmmTS;
  }

  // :69-103.  (0.2+0.25), 0.25/(0.2+0.25), 1.1*2.0*0.145 ... are REAL(4) constant expressions.
  void WearFactors(WearingFactors<R>& wearF)
  {
    const R Tph = settings.Tph;
    const float c_snow = 0.2f + 0.25f;
    wearF.SnowTran = R(static_cast<double>(c_snow)) * surf.SrfSnowmms;
    wearF.SnowTran = r_max(wearF.SnowTran, F4(0.01));
    if (surf.SrfSnowmms < F4(0.2)) wearF.SnowTran = wearF.SnowTran * R(3);
    const float s2i = 0.25f / (0.2f + 0.25f);
    condParam.Snow2IceFac = R(static_cast<double>(s2i));
    wearF.SnowTran = wearF.SnowTran * Tph;
    const float c_ice = 1.1f * 2.0f * 0.145f;
    wearF.IceWear = R(static_cast<double>(c_ice)) * surf.SrfIcemms;
    wearF.IceWear = r_max(wearF.IceWear, F4(0.01));
    wearF.IceWear = wearF.IceWear * Tph;
    const float c_ice2 = 1.1f * 2.0f * (4.0f * 0.290f);
    wearF.IceWear2 = R(static_cast<double>(c_ice2)) * surf.SrfIce2mms;
    wearF.IceWear2 = r_max(wearF.IceWear2, F4(0.01));
    wearF.IceWear2 = wearF.IceWear2 * Tph;
    const float c_dep = 0.5f * 2.0f * (4.0f * 0.290f);
    wearF.DepWear = R(static_cast<double>(c_dep)) * surf.SrfDepmms;
    wearF.DepWear = r_max(wearF.DepWear, F4(0.01));
    wearF.DepWear = wearF.DepWear * Tph;
    wearF.WatWear = F4(0.145) * surf.SrfWatmms;
    wearF.WatWear = r_max(wearF.WatWear, F4(0.06));
    wearF.WatWear = R(10) * wearF.WatWear * Tph;
  }

  // =================================== src/Storage.f90 ========================================

  // :33-84
  void WaterStorage(R MaxPormms, R& WatWear, R& SrfExtmms, R& SrfPormms)
  {
    const RoadCondParameters<R>& CP = condParam;
    if (surf.SrfSnowmms <= R(0.0) && surf.SrfIcemms <= R(0.0) && surf.SrfDepmms <= R(0.0) &&
        surf.TsurfAve > CP.TLimDew)
    {
      if (surf.SrfWatmms > MaxPormms)
        surf.SrfWatmms = surf.SrfWatmms - surf.EvapmmTS;
      else
        surf.SrfWatmms = surf.SrfWatmms - CP.PorEvaF * surf.EvapmmTS;
    }
    if (surf.WearSurf && surf.SrfWatmms > R(0.0))
    {
      if (surf.SrfWatmms < CP.WWearLim) WatWear = R(0.0);
      if (surf.SrfWatmms > CP.WWetLim)
        surf.SrfWatmms = surf.SrfWatmms - WatWear;
      else
        surf.SrfWatmms = surf.SrfWatmms - CP.DampWearF * WatWear;
    }
    if (surf.SrfWatmms < CP.MinWatmms) surf.SrfWatmms = R(0.0);
    if (surf.SrfWatmms > CP.MaxWatmms) surf.SrfWatmms = CP.MaxWatmms;
    SrfExtmms = r_max(surf.SrfWatmms - MaxPormms, R(0.));
    SrfPormms = r_min(surf.SrfWatmms, MaxPormms);
  }

  // :88-196
  void SnowStorage(R& SrfExtmms, R& Melted, R DTSecs, const WearingFactors<R>& wearF, R& SrfPormms,
                   R MaxPormms)
  {
    RoadCondParameters<R>& CP = condParam;
    R WatSnowRat = R(0.0);
    R RDummy = SrfExtmms + surf.SrfSnowmms;
    if (RDummy > F4(0.001))
      WatSnowRat = SrfExtmms / RDummy;
    else
      WatSnowRat = R(0.0);
    RDummy = surf.SrfSnowmms + surf.SrfIcemms;
    if (RDummy > F4(0.001))
      CP.SnowIceRat = surf.SrfSnowmms / RDummy;
    else
      CP.SnowIceRat = R(0.0);
    if (surf.SrfSnowmms > R(0.0))
    {
      if (WatSnowRat > CP.WetSnowFormR) atm.SnowType = SURFACE_SNOW_WET;
    }
    else
    {
      atm.SnowType = SURFACE_SNOW_DRY;
    }
    if (surf.SrfSnowmms > R(0.0))
    {
      if (surf.SrfDepmms > R(0.0))
      {
        surf.SrfIcemms = surf.SrfIcemms + surf.SrfDepmms;
        surf.SrfDepmms = R(0.0);
      }
    }
    if (surf.SrfSnowmms > R(0.0))
    {
      if (CP.forceSnowMelting)
      {
        surf.SrfWatmms = surf.SrfWatmms + surf.SrfSnowmms;
        surf.SrfSnowmms = R(0.0);
      }
      else if (surf.Q2Melt > R(0.0) && surf.TsurfAve >= CP.TLimMeltSnow)
      {
        Melted = (surf.Q2Melt * DTSecs) / (CP.WatMHeat * CP.WatDens);
        surf.SrfSnowmms = surf.SrfSnowmms - R(1000.) * Melted;
        surf.SrfWatmms = surf.SrfWatmms + R(1000.) * Melted;
      }
    }
    if (surf.WearSurf && surf.SrfSnowmms > R(0.0))
    {
      surf.SrfSnowmms = surf.SrfSnowmms - wearF.SnowTran;
      surf.SrfIcemms = surf.SrfIcemms + CP.Snow2IceFac * wearF.SnowTran;
      surf.SrfIce2mms = surf.SrfIce2mms + CP.Snow2IceFac * wearF.SnowTran;
    }
    if (surf.SrfSnowmms > R(0.0) && atm.SnowType == SURFACE_SNOW_WET)
    {
      if (WatSnowRat > CP.WetSnowMeltR)
      {
        surf.SrfWatmms = surf.SrfWatmms + surf.SrfSnowmms;
        surf.SrfSnowmms = R(0.0);
        atm.SnowType = SURFACE_SNOW_DRY;
      }
      if (surf.TsurfAve < CP.TLimFreeze)
      {
        surf.SrfIcemms = surf.SrfIcemms + surf.SrfSnowmms + surf.SrfWatmms;
        surf.SrfIce2mms = surf.SrfIce2mms + surf.SrfSnowmms + surf.SrfWatmms;
        atm.SnowType = SURFACE_SNOW_DRY;
        if (surf.SrfSnowmms > R(0.5)) CP.WetSnowFrozen = true;
        surf.SrfSnowmms = R(0.0);
        surf.SrfWatmms = R(0.0);
      }
    }
    SrfExtmms = r_max(surf.SrfWatmms - MaxPormms, R(0.));
    SrfPormms = r_min(surf.SrfWatmms, MaxPormms);
    if (surf.SrfSnowmms < CP.MinSnowmms) surf.SrfSnowmms = R(0.0);
    if (surf.SrfSnowmms > CP.MaxSnowmms) surf.SrfSnowmms = surf.SrfSnowmms - (CP.MaxSnowmms / R(2.));
  }

  // :199-267
  void IceStorage(R& Melted, R& SrfExtmms, R& SrfPormms, R MaxPormms, R DTSecs,
                  const WearingFactors<R>& wearF)
  {
    const RoadCondParameters<R>& CP = condParam;
    if (surf.TsurfAve < CP.TLimFreeze && surf.SrfWatmms > R(0.0))
    {
      surf.SrfIcemms = surf.SrfIcemms + surf.SrfWatmms;
      surf.SrfIce2mms = surf.SrfIce2mms + surf.SrfWatmms;
      surf.SrfWatmms = R(0.0);
    }
    if (surf.SrfSnowmms <= R(0.) && surf.SrfIcemms > R(0.))
    {
      if (CP.forceIceMelting)
      {
        surf.SrfWatmms = surf.SrfWatmms + surf.SrfIcemms;
        surf.SrfIcemms = R(0.0);
        surf.SrfIce2mms = R(0.0);
      }
      else if (surf.Q2Melt > R(0.0) && surf.TsurfAve >= CP.TLimMeltIce)
      {
        Melted = (surf.Q2Melt * DTSecs) / (CP.WatMHeat * CP.WatDens);
        surf.SrfIcemms = surf.SrfIcemms - R(1000.) * Melted;
        surf.SrfIce2mms = surf.SrfIce2mms - R(1000.) * Melted;
        surf.SrfWatmms = surf.SrfWatmms + R(1000.) * Melted;
      }
    }
    if (surf.WearSurf && surf.SrfIcemms > R(0.)) surf.SrfIcemms = surf.SrfIcemms - wearF.IceWear;
    if (surf.WearSurf && surf.SrfIce2mms > R(0.)) surf.SrfIce2mms = surf.SrfIce2mms - wearF.IceWear2;
    SrfExtmms = r_max(surf.SrfWatmms - MaxPormms, R(0.));
    SrfPormms = r_min(surf.SrfWatmms, MaxPormms);
    if (surf.SrfIcemms < CP.MinIcemms) surf.SrfIcemms = R(0.0);
    if (surf.SrfIcemms > CP.MaxIcemms) surf.SrfIcemms = CP.MaxIcemms;
    if (surf.SrfIce2mms < CP.MinIcemms) surf.SrfIce2mms = R(0.0);
    if (surf.SrfIce2mms > CP.MaxIcemms) surf.SrfIce2mms = CP.MaxIcemms;
  }

  // :271-314
  void DepositStorage(R DepWear, R& SrfExtmms, R& SrfPormms, R MaxPormms)
  {
    const RoadCondParameters<R>& CP = condParam;
    if (surf.EvapmmTS < R(0.0)) surf.SrfDepmms = surf.SrfDepmms - surf.EvapmmTS;
    if (surf.TsurfAve > CP.TLimMeltDep)
    {
      surf.SrfWatmms = surf.SrfWatmms + surf.SrfDepmms;
      surf.SrfDepmms = R(0.0);
    }
    if (surf.WearSurf && surf.SrfSnowmms <= R(0.0) && surf.SrfDepmms > R(0))
      surf.SrfDepmms = surf.SrfDepmms - DepWear;
    SrfExtmms = r_max(surf.SrfWatmms - MaxPormms, R(0.));
    SrfPormms = r_min(surf.SrfWatmms, MaxPormms);
    if (surf.SrfDepmms < CP.MinDepmms) surf.SrfDepmms = R(0.0);
    if (surf.SrfDepmms > CP.MaxDepmms)
    {
      surf.SrfWatmms = surf.SrfWatmms + (surf.SrfDepmms - CP.MaxDepmms);
      surf.SrfDepmms = CP.MaxDepmms;
    }
  }

  // :319-402
  void melting(bool inCouplingPhase, R TsurfObsLast, R depth)
  {
    const RoadCondParameters<R>& CP = condParam;
    if (surf.SrfSnowmms > R(0.0) || surf.SrfIcemms > R(0.0) || surf.SrfIce2mms > R(0.0))
    {
      do
      {
        if (!CP.CanMeltingChangeTemperature) break;
        if (ground.HStor <= F4(0.00001) || surf.TsurfAve <= surf.T4Melt || surf.Q2Melt <= R(0) ||
            (inCouplingPhase && TsurfObsLast < surf.T4Melt))
        {
          if (surf.TsurfAve < R(0.5))
          {
            surf.Q2Melt = R(0.0);
            break;
          }
          else if (surf.TsurfAve > R(2.0))
          {
            const R QAvail2 = ground.HS[1] * (ground.TmpNw[1] - surf.T4Melt);
            if (QAvail2 < surf.Q2Melt) surf.Q2Melt = QAvail2;
            break;
          }
        }
        const R QAvail = ground.HS[1] * (ground.TmpNw[1] - surf.T4Melt);
        if (surf.Q2Melt >= QAvail)
        {
          surf.Q2Melt = QAvail;
          ground.TmpNw[1] = surf.T4Melt + F4(0.01);
          ground.TmpNw[2] = surf.T4Melt + F4(0.01);
        }
        else
        {
          const R QLeftOver = QAvail - surf.Q2Melt;
          ground.TmpNw[1] = surf.T4Melt + (QLeftOver / ground.HS[1]);
          ground.TmpNw[2] = surf.T4Melt + F4(0.01);
        }
        if (depth >= R(0))
          surf.TsurfAve = getTempAtDepth(depth);
        else
          surf.TsurfAve = R(0.5) * (ground.TmpNw[1] + ground.TmpNw[2]);
      } while (false);
    }
    else
    {
      surf.Q2Melt = R(0.0);
    }
  }

  // :409-432
  void NewMeltFreezeHeat(R DTSecs)
  {
    const RoadCondParameters<R>& CP = condParam;
    surf.Q2Melt = R(0.0);
    if (surf.SrfSnowmms > R(0.0))
    {
      surf.Q2Melt = CP.WatMHeat * CP.WatDens * (surf.SrfSnowmms / R(1000.)) / DTSecs;
      surf.T4Melt = CP.TLimMeltSnow;
    }
    if (surf.SrfSnowmms <= R(0.0) && surf.SrfIcemms > R(0.0))
    {
      surf.Q2Melt = CP.WatMHeat * CP.WatDens * (surf.SrfIcemms / R(1000.)) / DTSecs;
      surf.T4Melt = CP.TLimMeltIce;
    }
    if (surf.Q2Melt < R(0.0)) surf.Q2Melt = R(0.0);
  }

  // src/Cond.f90:9-65
  void RoadCond(R MaxPormms, WearingFactors<R>& wearF)
  {
    RoadCondParameters<R>& CP = condParam;
    R SrfExtmms = R(0), SrfPormms = R(0);
    R Melted = R(0.0);
    CP.SnowIceRat = R(0.0);
    atm.SnowType = SURFACE_SNOW_DRY;
    if (surf.VeryCold && surf.TsurfAve > CP.TLimColdH) surf.VeryCold = false;
    if (!surf.VeryCold && surf.TsurfAve < CP.TLimColdL) surf.VeryCold = true;
    WaterStorage(MaxPormms, wearF.WatWear, SrfExtmms, SrfPormms);
    SnowStorage(SrfExtmms, Melted, settings.DTSecs, wearF, SrfPormms, MaxPormms);
    IceStorage(Melted, SrfExtmms, SrfPormms, MaxPormms, settings.DTSecs, wearF);
    DepositStorage(wearF.DepWear, SrfExtmms, SrfPormms, MaxPormms);
    if (surf.SrfWatmms < CP.MinWatmms) surf.SrfWatmms = R(0.0);
    if (surf.SrfWatmms > CP.MaxWatmms) surf.SrfWatmms = CP.MaxWatmms;
    NewMeltFreezeHeat(settings.DTSecs);
  }

  // src/Cond.f90:105-139
  void CalcAlbedo()
  {
    const RoadCondParameters<R>& CP = condParam;
    if (surf.WearSurf)
    {
      R IceSum = R(0.5) * (surf.SrfIcemms + surf.SrfIce2mms) + surf.SrfDepmms;
      if (IceSum < R(0.0)) IceSum = R(0.0);
      const R IceMax = R(1.5);
      ground.Albedo = CP.AlbDry;
      if (surf.SrfSnowmms > F4(0.01) && surf.SrfSnowmms > surf.SrfIcemms)
      {
        ground.Albedo = CP.AlbSnow;
      }
      else if (surf.SrfIcemms > F4(0.01) || surf.SrfDepmms > F4(0.01))
      {
        if (IceSum < IceMax)
          ground.Albedo = CP.AlbDry + (IceSum / IceMax) * (CP.AlbSnow - CP.AlbDry);
        else
          ground.Albedo = CP.AlbSnow;
      }
    }
  }

  // =================================== src/SunPosition.f90 ====================================

  // :196-260.  REAL(int) is REAL(4); the day fraction is summed in single precision.
  R JulianEphemerisDay(int idx) const
  {
    const int mmyr = modelInput.year[idx - 1], mmmon = modelInput.month[idx - 1],
              mmday = modelInput.day[idx - 1], mmhr = modelInput.hour[idx - 1],
              mmmin = modelInput.minute[idx - 1], mmsec = modelInput.second[idx - 1];
    const R Dyr = R(365.25);
    R yr, mo;
    if (mmmon <= 2)
    {
      yr = R(static_cast<double>(static_cast<float>(mmyr - 1)));
      mo = R(static_cast<double>(static_cast<float>(mmmon + 12)));
    }
    else
    {
      yr = R(static_cast<double>(static_cast<float>(mmyr)));
      mo = R(static_cast<double>(static_cast<float>(mmmon)));
    }
    const float dayf = static_cast<float>(mmday) + static_cast<float>(mmhr) / 24.f +
                       static_cast<float>(mmmin) / (24.f * 60.f) +
                       static_cast<float>(mmsec) / (24.f * 60.f * 60.f);
    const R day = R(static_cast<double>(dayf));
    const R A = r_aint(yr / R(100.));
    const R B = R(2.) - A + r_aint(A / R(4.));
    return r_aint(Dyr * (yr + R(4716))) + r_aint(F4(30.6001) * (mo + R(1.))) + day + B - R(1.5245e3);
  }

  // :20-194
  void calcElevationAzimuth(R JDE, R lat, R lon, R& elevation_angle, R& azimuth_angle)
  {
    const R pi = R(4) * R(std::atan(1.0));
    const R Dyr = R(365.25);
    const R T = (JDE - R(2451545.0)) / (Dyr * R(100.));
    R ml = F4(280.46645) + F4(36000.76983) * T + F4(0.0003032) * T * T;
    if (ml < R(0.)) ml = ml - R(360.) * (r_aint(ml / R(360.)) - R(1.));
    if (ml > R(360.)) ml = ml - R(360.) * r_aint(ml / R(360.));
    R ma = F4(357.52910) + F4(35999.05030) * T - F4(0.0001559) * T * T - F4(0.00000048) * T * T * T;
    if (ma < R(0.)) ma = ma - R(360.) * (r_aint(ma / R(360.)) - R(1.));
    if (ma > R(360.)) ma = ma - R(360.) * r_aint(ma / R(360.));
    const R ecc = F4(0.016708617) - F4(0.000042037) * T - F4(0.0000001236) * T * T;
    (void)ecc;
    const R sunc = (F4(1.913600) - F4(0.004817) * T - F4(0.000014) * T * T) * r_sin(ma * pi / R(180.)) +
                   (F4(0.019993) - F4(0.000101) * T) * r_sin(R(2.) * ma * pi / R(180.)) +
                   F4(0.000290) * r_sin(R(3.) * ma * pi / R(180.));
    R al = ml + sunc - F4(0.00569) - F4(0.00478) * r_sin((F4(125.04) - F4(1934.136) * T) * pi / R(180.));
    al = al * pi / R(180.);
    const R tilt =
        F4(23.43929111) - F4(0.013004166) * T - F4(0.001638888) * T * T + F4(0.005036111) * T * T * T;
    R eps = tilt + F4(0.00256) * r_cos((F4(125.04) - F4(1934.136) * T) * pi / R(180.));
    eps = eps * pi / R(180.);
    R ra = r_atan2(r_cos(eps) * r_sin(al), r_cos(al));
    if (ra < R(0.)) ra = ra - R(2.) * pi * (r_aint(ra / (R(2.) * pi)) - R(1.));
    if (ra > R(2.) * pi) ra = ra - R(2.) * pi * r_aint(ra / (R(2.) * pi));
    const R declination = r_asin(r_sin(eps) * r_sin(al));
    R stG = F4(280.46061837) + F4(360.98564736629) * (JDE - R(2451545.0)) + F4(0.000387933) * T * T -
            T * T * T / R(38710000.);
    if (stG < R(0.)) stG = stG - R(360.) * (r_aint(stG / R(360.)) - R(1.));
    if (stG > R(360.)) stG = stG - R(360.) * r_aint(stG / R(360.));
    stG = stG * pi / R(180.);
    const R cos_declination = r_cos(declination);
    const R sin_declination = r_sin(declination);
    const R lat_radians = pi * lat / R(180.);
    const R sin_lat = r_sin(lat_radians);
    const R cos_lat = r_cos(lat_radians);
    const R cos_dec_lat = cos_declination * cos_lat;
    const R sin_dec_lat = sin_declination * sin_lat;
    R hour_angle_corr = (stG + lon * pi / R(180.) - ra);
    if (ra < R(0.))
      hour_angle_corr = hour_angle_corr - R(2.) * pi * (r_aint(hour_angle_corr / (R(2.) * pi)) - R(1.));
    if (ra > R(2.) * pi)
      hour_angle_corr = hour_angle_corr - R(2.) * pi * r_aint(hour_angle_corr / (R(2.) * pi));
    const R cosah = r_cos(hour_angle_corr);
    R cos_elev = sin_dec_lat + cos_dec_lat * cosah;
    R chi;
    if (cos_elev >= R(1.0) && cos_elev < F4(1.001))
    {
      cos_elev = R(1.0);
      chi = R(0.);
    }
    else if (cos_elev >= F4(1.001))
    {
      diag.solar_stop = true;  // Fortran: stop
      chi = R(0.);
    }
    else if (cos_elev > F4(-1.001) && cos_elev <= R(-1.0))
    {
      cos_elev = R(-1.0);
      chi = pi;
    }
    else
    {
      chi = r_acos(cos_elev);
    }
    elevation_angle = R(90.0) - chi * (R(180.) / pi);
    if (hour_angle_corr < R(0.))
      hour_angle_corr = R(2) * pi + hour_angle_corr;
    else if (hour_angle_corr > R(2) * pi)
      hour_angle_corr = hour_angle_corr - R(2) * pi;
    if (elevation_angle > R(0))
    {
      const R cosele = r_cos((pi / R(2.0)) - chi);
      R precos = R(0.0);
      if (cosele >= F4(-0.0001) && cosele < F4(0.0001))
      {
        azimuth_angle = F4(-9999.9);
      }
      else
      {
        precos = (sin_declination * cos_lat - cos_declination * sin_lat * cosah) / cosele;
        if (precos >= R(1.0) && precos < F4(1.001))
        {
          precos = R(1.0);
          azimuth_angle = R(0.0);
        }
        else if (precos >= F4(1.001))
        {
          diag.solar_stop = true;  // Fortran: stop
          azimuth_angle = R(0.0);
        }
        else if (precos > F4(-1.001) && precos <= R(-1.0))
        {
          precos = R(-1.0);
          azimuth_angle = pi;
        }
        else
        {
          azimuth_angle = r_acos(precos);
        }
      }
      if (hour_angle_corr < pi) azimuth_angle = R(2) * pi - azimuth_angle;
      azimuth_angle = azimuth_angle * (R(180.) / pi);
    }
    else
    {
      azimuth_angle = F4(-9999.9);
      elevation_angle = F4(-9999.9);
    }
  }

  // :4-17
  void SunPosition(const LocalParameters& lp, int i, R& elevation_angle, R& azimuth_angle)
  {
    const R JDE = JulianEphemerisDay(i);
    calcElevationAzimuth(JDE, R(lp.lat), R(lp.lon), elevation_angle, azimuth_angle);
  }

  // =================================== src/ModRadiation.f90:7-73 ==============================
  void ModRadiationBySurroundings(const InputParameters& ip, const LocalParameters& lp, int i)
  {
    const int k = i - 1;
    R dif_SW = R(modelInput.SW[k]) - R(modelInput.SW_dir[k]);
    const R LW_surroundings = R(modelInput.LW_net[k]) - R(modelInput.LW[k]);
    R sun_elevation, sun_azim;
    SunPosition(lp, i, sun_elevation, sun_azim);
    R horizon_in_sun_dir = R(0.);
    long azim_idx = std::lround(r_val(sun_azim));  // NINT
    if (azim_idx == 360) azim_idx = 0;
    // The reference reads local_horizons(-9999) when the sun is down (result unused); guard it.
    if (azim_idx >= 0 && azim_idx < 360) horizon_in_sun_dir = R(modelInput.local_horizons[azim_idx]);
    R shadow_fac;
    if (horizon_in_sun_dir > sun_elevation)
      shadow_fac = R(0.0);
    else
      shadow_fac = R(1.0);
    if (sun_elevation > R(0.0))
    {
      modelInput.SW_dir[k] = r_val(R(modelInput.SW_dir[k]) * shadow_fac);
      const R SW_ref = R(ip.Albedo_surroundings) * R(modelInput.SW_dir[k]) +
                       R(ip.Albedo_surroundings) * dif_SW;
      dif_SW = R(lp.sky_view) * dif_SW + (R(1.0) - R(lp.sky_view)) * SW_ref;
      modelInput.SW[k] = r_val(dif_SW + R(modelInput.SW_dir[k]));
    }
    modelInput.LW[k] =
        r_val(R(lp.sky_view) * R(modelInput.LW[k]) + (R(1.0) - R(lp.sky_view)) * (-LW_surroundings));
  }

  // =================================== src/BalanceModel.f90 ===================================

  // :354-387
  void SetDayDependendVariables(int inputIdx)
  {
    const int shour = modelInput.hour[inputIdx - 1];
    if (R(shour) >= settings.NightOn || R(shour) <= settings.NightOff)
    {
      atm.CalmLim = settings.CalmLimNgt;
      surf.TrfFric = settings.TrfFricNgt;
    }
    else
    {
      atm.CalmLim = settings.CalmLimDay;
      surf.TrfFric = settings.TrFfricDay;
    }
    if (atm.VZ < atm.CalmLim) atm.VZ = atm.CalmLim;
  }

  // :282-307
  static void CalcRNet(R Emiss, R SB_Const, R TSurfAve, R Albedo, R SW, R LW, R& RNet, R SwRadCof,
                       R LWRadCof)
  {
    const R TsurfK = TSurfAve + F4(273.15);
    const R TsurfK2 = TsurfK * TsurfK;
    const R RBB = Emiss * SB_Const * (TsurfK2 * TsurfK2);
    RNet = (R(1.) - Albedo) * SW * SwRadCof + Emiss * LW * LWRadCof - RBB;
  }

  // :90-129
  void calcProfile()
  {
    const int n = settings.NLayers;
    const R DTSecs = settings.DTSecs;
    std::vector<R> GFlux(n + 2, R(0));
    atm.SensibleHeatFlux = atm.BLCond * (ground.Tmp[0] - ground.Tmp[1]);
    GFlux[0] = atm.RNet - atm.LE_Flux + surf.TrfFric + atm.SensibleHeatFlux;
    ground.TmpNw = ground.Tmp;
    for (int j = 1; j <= n; ++j) GFlux[j] = ground.condDZ[j] * (ground.Tmp[j + 1] - ground.Tmp[j]);
    ground.GroundFlux = GFlux[3];
    for (int j = 1; j <= n; ++j)
      ground.TmpNw[j] = ground.Tmp[j] + DTSecs * (ground.capDZ[j] * (GFlux[j] - GFlux[j - 1]));
  }

  // :311-322
  void calcHStor()
  {
    const R T1Ave = (ground.Tmp[1] + R(3.) * ground.Tmp[2]) / R(4.);
    const R TN1Ave = (ground.TmpNw[1] + R(3.) * ground.TmpNw[2]) / R(4.);
    ground.HStor = ground.HS[1] * (TN1Ave - T1Ave);
  }

  // :7-86
  void BalanceModelOneStep(R SWi, R LWi, int inputIdx)
  {
    SetDayDependendVariables(inputIdx);
    CalcBLCondAndLE();
    CalcRNet(phy.Emiss, phy.SB_const, surf.TsurfAve, ground.Albedo, SWi, LWi, atm.RNet,
             coupling.SWRadCof, coupling.LWRadCof);
    CalcHCapHCond();
    calcCapDZCondDZ();
    calcProfile();
    calcHStor();
    R depth;
    if (settings.tsurfOutputDepth >= R(0.0))
      depth = settings.tsurfOutputDepth;
    else
      depth = R(modelInput.depth[inputIdx - 1]);
    melting(coupling.inCouplingPhase, coupling.lastTsurfObs, depth);
    ground.Tmp = ground.TmpNw;
    if (depth >= R(0))
      surf.TsurfAve = getTempAtDepth(depth);
    else
      surf.TsurfAve = (ground.Tmp[1] + ground.Tmp[2]) / R(2.0);
  }

  // =================================== src/Coupling.f90 =======================================

  // :172-210 (SrfIcemms is never saved: the reference assigns srfIce2mmsSave twice)
  void saveDataForCoupling(int datai)
  {
    const int couplingLen = coupling.couplingEndI[1] - coupling.couplingStartI[1] + 1;
    coupling.saveDatai = datai;
    coupling.TSurfAveSave = surf.TsurfAve;
    coupling.SrfWatmmsSave = surf.SrfWatmms;
    coupling.SrfIce2mmsSave = surf.SrfIce2mms;
    coupling.SrfIce2mmsSave = surf.SrfIce2mms;
    coupling.SrfDepmmsSave = surf.SrfDepmms;
    coupling.SrfSnowmmsSave = surf.SrfSnowmms;
    coupling.AlbedoSave = ground.Albedo;
    coupling.VeryColdSave = surf.VeryCold;
    for (int i = 0; i <= settings.NLayers + 1; ++i) coupling.TmpSave[i] = ground.Tmp[i];
    for (int i = 1; i <= couplingLen; ++i)
    {
      const int k = coupling.couplingStartI[1] + i - 1 - 1;
      coupling.SWSave[i] = R(modelInput.SW[k]);
      coupling.SWDirSave[i] = R(modelInput.SW_dir[k]);
      coupling.LWSave[i] = R(modelInput.LW[k]);
    }
  }

  // :213-255
  void uploadDataForCoupling(int& datai)
  {
    const int couplingLen = coupling.couplingEndI[1] - coupling.couplingStartI[1] + 1;
    datai = coupling.saveDatai;
    surf.TsurfAve = coupling.TSurfAveSave;
    surf.SrfWatmms = coupling.SrfWatmmsSave;
    surf.SrfIce2mms = coupling.SrfIce2mmsSave;
    surf.SrfIce2mms = coupling.SrfIce2mmsSave;
    surf.SrfDepmms = coupling.SrfDepmmsSave;
    surf.SrfSnowmms = coupling.SrfSnowmmsSave;
    ground.Albedo = coupling.AlbedoSave;
    surf.VeryCold = coupling.VeryColdSave;
    for (int i = 0; i <= settings.NLayers + 1; ++i) ground.Tmp[i] = coupling.TmpSave[i];
    for (int i = 1; i <= couplingLen; ++i)
    {
      const int k = coupling.couplingStartI[1] + i - 1 - 1;
      modelInput.SW[k] = r_val(coupling.SWSave[i]);
      modelInput.SW_dir[k] = r_val(coupling.SWDirSave[i]);
      modelInput.LW[k] = r_val(coupling.LWSave[i]);
    }
  }

  // :259-289
  void snowIceCheck(R LastTsurfObs)
  {
    const RoadCondParameters<R>& CP = condParam;
    if (LastTsurfObs > CP.TLimMeltSnow && surf.SrfSnowmms > R(0.00))
    {
      surf.SrfWatmms = surf.SrfWatmms + surf.SrfSnowmms;
      surf.SrfSnowmms = R(0.00);
    }
    if (LastTsurfObs > CP.TLimMeltIce && surf.SrfIcemms > R(0.00))
    {
      surf.SrfWatmms = surf.SrfWatmms + surf.SrfIcemms;
      surf.SrfIcemms = R(0.00);
    }
    if (LastTsurfObs > CP.TLimMeltIce && surf.SrfIce2mms > R(0.00))
    {
      surf.SrfIce2mms = R(0.00);
    }
    if (LastTsurfObs > CP.TLimMeltDep && surf.SrfDepmms > R(0.00))
    {
      surf.SrfWatmms = surf.SrfWatmms + surf.SrfDepmms;
      surf.SrfDepmms = R(0.00);
    }
  }

  // :10-96
  void CouplingOperations1(int& i, const LocalParameters& lp)
  {
    const R DTs = settings.DTSecs;
    const int N = coupling.CoupPhaseN;
    coupling.inCouplingPhase = false;
    if (i >= coupling.couplingStartI[N] && i <= coupling.couplingEndI[N])
      coupling.inCouplingPhase = true;
    if (i == coupling.couplingStartI[N] && coupling.Coupling_iterations == 0)
    {
      saveDataForCoupling(i);
      coupling.SWRadCof = R(1.0);
      coupling.LWRadCof = R(1.0);
      coupling.SW_correction = R(0.0);
      coupling.LW_correction = R(0.0);
    }
    if (coupling.start_coupling_again)
    {
      uploadDataForCoupling(i);
      ++diag.coupling_restarts;
      coupling.start_coupling_again = false;
      if (R(modelInput.SW[i - 1]) > R(modelInput.LW[i - 1]) && !sky_view_active(lp))
      {
        coupling.SWRadCof = coupling.RadCoeff;
        coupling.LWRadCof = R(1.0);
      }
      else
      {
        coupling.SWRadCof = R(1.0);
        coupling.LWRadCof = coupling.RadCoeff;
      }
    }
    if (i > coupling.couplingEndI[N])
    {
      coupling.SWRadCof =
          R(1.0) + coupling.SW_correction *
                       r_exp(-((DTs * R(i)) - (DTs * R(coupling.couplingEndI[N]))) /
                             settings.couplingEffectReduction);
      coupling.LWRadCof =
          R(1.0) + coupling.LW_correction *
                       r_exp(-((DTs * R(i)) - (DTs * R(coupling.couplingEndI[N]))) /
                             settings.couplingEffectReduction);
    }
    if (coupling.inCouplingPhase) snowIceCheck(coupling.lastTsurfObs);
  }

  // :292-481
  void Coupling_control(R& TsurfAve)
  {
    CouplingVariables<R>& c = coupling;
    c.start_coupling_again = false;
    TsurfAve = TsurfAve + F4(273.16);
    c.lastTsurfObs = c.lastTsurfObs + F4(273.16);
    if (!c.Coupling_failed)
    {
      if (c.Coupling_iterations == 0) c.Tsurf_end_coup1 = TsurfAve;
      if (c.Coupling_iterations == 25)
      {
        if (r_abs(c.Tsurf_end_coup1 - c.lastTsurfObs) < r_abs(TsurfAve - c.lastTsurfObs))
          c.start_coupling_again = true;
        c.SWRadCof = R(1.0);
        c.LWRadCof = R(1.0);
        c.SW_correction = R(0.0);
        c.LW_correction = R(0.0);
        c.RadCoeff = R(1.0);
        c.Coupling_failed = true;
      }
      else if (c.lastTsurfObs < R(-100))
      {
        c.SWRadCof = R(1.0);
        c.LWRadCof = R(1.0);
        c.SW_correction = R(0.0);
        c.LW_correction = R(0.0);
        c.RadCoeff = R(1.0);
        c.Coupling_failed = true;
        c.start_coupling_again = true;
      }
      else if (TsurfAve < R(170.0) || TsurfAve > R(400.0) || c.Coupling_failed)
      {
        c.SWRadCof = R(1.0);
        c.LWRadCof = R(1.0);
        c.SW_correction = R(0.0);
        c.LW_correction = R(0.0);
        c.Coupling_failed = true;
        c.start_coupling_again = true;
        c.RadCoeff = R(1.0);
      }
      else if (TsurfAve - c.lastTsurfObs > F4(0.1))
      {
        if (c.TsurfNearestAbove < R(-100))
        {
          c.TsurfNearestAbove = TsurfAve;
          c.RadCoefNearestAbove = c.RadCoeff;
        }
        else if (c.TsurfNearestAbove - c.lastTsurfObs > TsurfAve - c.lastTsurfObs)
        {
          c.TsurfNearestAbove = TsurfAve;
          c.RadCoefNearestAbove = c.RadCoeff;
        }
        c.start_coupling_again = true;
        if (c.TsurfNearestAbove > R(-100) && c.TsurfNearestBelow > R(-100))
        {
          const R TDifAbove = c.TsurfNearestAbove - c.lastTsurfObs;
          const R TDifBelow = c.lastTsurfObs - c.TsurfNearestBelow;
          c.RadCoeff = c.RadCoefNearestAbove - TDifAbove / (TDifAbove + TDifBelow) *
                                                   (c.RadCoefNearestAbove - c.RadCoefNearestBelow);
        }
        else
        {
          c.RadCoeff = R(0.5) * c.RadCoeff;
        }
        if (r_abs(c.RadCoeff - c.RadCoeffPrevious) < F4(0.00005))
        {
          c.TsurfNearestAbove = R(-9999);
          c.TsurfNearestBelow = R(-9999);
        }
        if (c.RadCoeff < F4(0.01))
        {
          if (diag.verbose) std::printf(" coupling coefficient too small, coupling failed\n");
          c.RadCoeff = R(1.0);
          c.Coupling_failed = true;
          c.SWRadCof = R(1.0);
          c.LWRadCof = R(1.0);
          c.SW_correction = R(0.0);
          c.LW_correction = R(0.0);
        }
        c.RadCoeffPrevious = c.RadCoeff;
      }
      else if (c.lastTsurfObs - TsurfAve > F4(0.1))
      {
        if (c.TsurfNearestBelow < R(-100))
        {
          c.TsurfNearestBelow = TsurfAve;
          c.RadCoefNearestBelow = c.RadCoeff;
        }
        else if (c.TsurfNearestBelow - c.lastTsurfObs < TsurfAve - c.lastTsurfObs)
        {
          c.TsurfNearestBelow = TsurfAve;
          c.RadCoefNearestBelow = c.RadCoeff;
        }
        c.start_coupling_again = true;
        if (c.TsurfNearestAbove > R(-100) && c.TsurfNearestBelow > R(-100))
        {
          const R TDifAbove = c.TsurfNearestAbove - c.lastTsurfObs;
          const R TDifBelow = c.lastTsurfObs - c.TsurfNearestBelow;
          c.RadCoeff = c.RadCoefNearestAbove - TDifAbove / (TDifAbove + TDifBelow) *
                                                   (c.RadCoefNearestAbove - c.RadCoefNearestBelow);
        }
        else
        {
          c.RadCoeff = R(2.0) * c.RadCoeff;
        }
        if (r_abs(c.RadCoeff - c.RadCoeffPrevious) < F4(0.00005))
        {
          c.TsurfNearestAbove = R(-9999);
          c.TsurfNearestBelow = R(-9999);
        }
        c.RadCoeffPrevious = c.RadCoeff;
      }
      else
      {
        if (c.RadCoeff > R(3.0))
        {
          if (diag.verbose) std::printf(" coupling coefficient too big, coupling failed\n");
          c.Coupling_failed = true;
          c.RadCoeff = R(1.0);
          c.SWRadCof = R(1.0);
          c.LWRadCof = R(1.0);
          c.SW_correction = R(0.0);
          c.LW_correction = R(0.0);
        }
        c.SW_correction = c.SWRadCof - R(1.0);
        c.LW_correction = c.LWRadCof - R(1.0);
        c.Coupling_failed = false;
        c.Coupling_iterations = -1;
        c.TsurfNearestAbove = R(-9999.0);
        c.TsurfNearestBelow = R(-9999.0);
        c.RadCoeff = R(1.0);
        c.RadCoefNearestAbove = R(-9999.0);
        c.RadCoefNearestBelow = R(-9999.0);
        c.RadCoeffPrevious = R(1.0);
        if (c.CoupPhaseN < c.NObs)
        {
          c.CoupPhaseN = c.CoupPhaseN + 1;
          c.lastTsurfObs = c.obsTsurf[c.CoupPhaseN] + F4(273.16);
        }
      }
    }
    TsurfAve = TsurfAve - F4(273.16);
    c.lastTsurfObs = c.lastTsurfObs - F4(273.16);
  }

  // :121-141
  void CouplingOperations2(R& TSurfAve)
  {
    if (coupling.Coupling_iterations == 0) coupling.Tsurf_end_coup1 = TSurfAve;
    Coupling_control(TSurfAve);
    coupling.Coupling_iterations = coupling.Coupling_iterations + 1;
  }

  // :98-118
  void CheckEndCoupling(int i)
  {
    if (settings.use_coupling && i == coupling.couplingEndI[coupling.CoupPhaseN] &&
        !coupling.Coupling_failed)
    {
      CouplingOperations2(surf.TsurfAve);
    }
  }

  // ============================ examples/example1/src/Simulation.f90 ==========================

  // :120-172
  void roadModelOneStep(int input_idxI, const InputParameters& ip, const LocalParameters& lp)
  {
    WearingFactors<R> wearF;
    double* tr = diag.trace ? diag.trace + static_cast<size_t>(input_idxI - 1) * TRACE_N : nullptr;
    if (tr)
    {
      tr[16] = -static_cast<double>(diag.bl_iterations);
      tr[17] = -static_cast<double>(diag.bl_unstable);
      tr[0] = r_val(surf.TsurfAve);
      tr[1] = r_val(atm.Tair);
      tr[2] = r_val(atm.PrecInTStep);
      tr[3] = r_val(surf.Q2Melt);
    }
    PrecipitationToStorage(modelInput.PrecPhase[input_idxI - 1]);
    if (sky_view_active(lp)) ModRadiationBySurroundings(ip, lp, input_idxI);
    BalanceModelOneStep(R(modelInput.SW[input_idxI - 1]), R(modelInput.LW[input_idxI - 1]),
                        input_idxI);
    if (tr)
    {
      tr[4] = r_val(atm.RainmmTS);
      tr[5] = r_val(atm.SnowmmTS);
      tr[6] = r_val(surf.SrfSnowmms);
      tr[7] = r_val(surf.SrfWatmms);
      tr[8] = r_val(surf.SrfIcemms);
      tr[9] = r_val(surf.Q2Melt);
      tr[10] = r_val(ground.Tmp[1]);
      tr[11] = r_val(ground.Tmp[2]);
      tr[12] = r_val(ground.HStor);
      tr[13] = r_val(atm.BLCond);
      tr[14] = r_val(atm.LE_Flux);
      tr[15] = r_val(surf.EvapmmTS);
      tr[16] += static_cast<double>(diag.bl_iterations);  // boundary-layer iterations of this step
      tr[17] += static_cast<double>(diag.bl_unstable);    // ... of which took the unstable branch
    }
    WearFactors(wearF);
    RoadCond(phy.MaxPormms, wearF);
    CalcAlbedo();
    ++diag.executed_steps;
  }

  // :4-117
  void runsimulation(OutputPointers& outP, const InputPointers& inP, const InputSettings& inS,
                     const InputParameters& ip, const LocalParameters& lp)
  {
    ConnectFortran2Carrays(inP, outP);
    Initialization(inS, ip, lp);
    diag.coupling_used = settings.use_coupling;
    int i = 1;
    while (i < settings.SimLen && !settings.simulation_failed)
    {
      CheckValues(i, lp);
      if (settings.use_coupling) CouplingOperations1(i, lp);
      SetCurrentValues(i);
      if (settings.use_relaxation) RelaxationOperations(i);
      roadModelOneStep(i, ip, lp);
      SaveOutput(i);
      CheckEndCoupling(i);
      i = i + 1;
    }
    if (!settings.simulation_failed)
    {
      lastValues();
      roadModelOneStep(settings.SimLen, ip, lp);
      SaveOutput(i);
    }
    diag.coupling_failed = coupling.Coupling_failed;
  }
};

#undef F4

}  // namespace rs_oracle
