// rs_model.h -- per-run constants of the step kernel (one copy in __constant__ memory per device).
//
// Everything here is identical for all points of a run: the 69 InputParameters, the settings and
// what the reference derives from them once per point in Initialization
// (src/Initialization.f90:181-235,310-358; src/BalanceModel.f90:132-186,254-279).  The reference
// recomputes these for every point; here they are computed once on the host and broadcast.
#pragma once

#include "../../include/roadsurf_b200.h"

#define RS_MAX_LAYERS 32
#define RS_MIN_LAYERS 4   // layers 1-4 / WCont(1:4) / GFlux(3) are touched unconditionally

struct RsModel
{
  // ---- settings (src/InputSettings.f90.inc) ----
  int nlayers;
  int use_coupling;       // int2Logical(inSettings%use_coupling)
  int use_relaxation;
  int force_tsurf;
  int sim_len_hint;       // informational
  int coupling_span;      // int(coupling_minutes*60/DTSecs)          src/Coupling.f90:517
  double coupling_span_real;  // coupling_minutes*60/DTSecs           src/Coupling.f90:512
  double DT;              // DTSecs
  double Tph;             // DTSecs/3600.0                            src/Initialization.f90:91
  double tsurfOutputDepth;
  double couplingEffectReduction;

  // ---- day/night traffic (src/BalanceModel.f90:354-387) ----
  double NightOn, NightOff, CalmLimDay, CalmLimNgt, TrfFricNgt, TrFfricDay;

  // ---- physical parameters (src/Initialization.f90:310-358) ----
  double Grav, SB_Const, VK_Const, LVap, LFus, ZRefT, Emiss, Albedo0, Albedo_surroundings, MaxPormms;
  double logMom, logHeat, logCond, logUstar;
  double TClimG, AZ, Omega, DampDpth;
  double dry1, dry2;      // (1-Poro1)*vsh1, (1-Poro2)*vsh2           src/BalanceModel.f90:232-236

  // ---- road condition parameters (src/Initialization.f90:479-557) ----
  double WatDens, WatMHeat, PorEvaF, DampWearF;
  double TLimFreeze, TLimMeltSnow, TLimMeltIce, TLimMeltDep, TLimDew, T4Melt0;
  double WetSnowFormR, WetSnowMeltR, PLimSnow, PLimRain;
  double MinPrecmm, MinWatmms, MinSnowmms, MinDepmms, MinIcemms;
  double MaxSnowmms, MaxDepmms, MaxIcemms, MaxWatmms;
  double AlbDry, AlbSnow, MissValI, WWetLim, WWearLim;

  // ---- ground geometry and conductivities, Fortran index = array index ----
  double ZDpth[RS_MAX_LAYERS + 2];   // 1..N+1                        src/Initialization.f90:217-235
  double DyC[RS_MAX_LAYERS + 2];     // 1..N                          src/Initialization.f90:191-195
  double condDZ[RS_MAX_LAYERS + 2];  // -(CC/DyK), 1..N               src/BalanceModel.f90:144,150
  double WCont[RS_MAX_LAYERS + 2];   // 1..N                          src/Initialization.f90:207-213
  double hs1_dz;                     // ZDpth(2)-ZDpth(1)             src/BalanceModel.f90:240-241
  double two_dt;                     // 2.0*DTSecs
  // Correctly rounded reciprocals of run constants: x / c is evaluated as q = x*rc,
  // q + fma(-q, c, x)*rc, which equals the IEEE quotient (see div_const in rs_kernel.cu).
  double inv_two_dt, inv_DT, inv_3600, inv_1000, inv_3364, inv_1p5, inv_4h, inv_CER;
  double WatMHeatDens, inv_WatMHeatDens;   // WatMHeat*WatDens and its reciprocal
  // fixed output depth (tsurfOutputDepth >= 0): 0 = mean of layers 1,2; 1 = Tmp(1);
  // 2 = Tmp(N+1); 3 = interpolate between depth_idx and depth_idx+1   src/BalanceModel.f90:390-417
  int depth_mode;
  int depth_idx;
};

// Derives an RsModel from the caller's settings and parameters.  Returns 0 on success, a negative
// RS_ERR_* code (message in `err`) otherwise.
int rs_build_model(const InputSettings* settings, const InputParameters* params, RsModel* model,
                   char* err, int errlen);
