// rs_model.cpp -- host-side derivation of the per-run constants (see rs_model.h).
//
// Fortran un-suffixed real literals are REAL(4): they are written here as float literals widened
// to double, and literal-only sub-expressions are evaluated in float, as gfortran does.
#include "rs_model.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace
{
inline double f4(float x) { return static_cast<double>(x); }

// REAL(4) ** INTEGER by square-and-multiply, as libgcc's __powisf2 evaluates it.
float powi_f32(float x, int n)
{
  unsigned m = static_cast<unsigned>(n < 0 ? -n : n);
  float y = (m & 1u) ? x : 1.0f;
  while (m >>= 1)
  {
    x *= x;
    if (m & 1u) y *= x;
  }
  return n < 0 ? 1.0f / y : y;
}

// Campbell soil heat conductivity for one water content (src/BalanceModel.f90:158-186,254-279).
double campbell_cc(double rhoB, double silt, double wcont)
{
  const double A = f4(0.65f) - f4(0.78f) * rhoB + f4(0.60f) * rhoB * rhoB;
  const double B = f4(1.06f) * rhoB;
  const double Cc = (silt > f4(0.00001f)) ? 1.0 + f4(2.6f) / std::sqrt(silt) : 0.0;
  const double D = f4(0.03f) + f4(0.1f) * rhoB * rhoB;
  const double E = 4.0;
  return A + B * wcont - (A - D) * std::exp(-std::pow(Cc * wcont, E));
}
}  // namespace

extern "C" void roadsurf_default_parameters(InputParameters* p, double DTSecs)
{
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  // examples/example1/src/InputParameters.h:18-110
  p->NightOn = 19.0;  p->NightOff = 4.0;  p->CalmLimDay = 1.5;  p->CalmLimNgt = 0.4;
  p->TrfFricNgt = 5.0;  p->TrFfricDay = 10.0;
  p->Grav = 9.81;  p->SB_Const = 5.67e-8;  p->VK_Const = 0.4;  p->LVap = 2.452e6;  p->LFus = 0.334e6;
  p->WatDens = 999.87;  p->SnowDens = 100.0;  p->IceDens = 920.0;  p->DepDens = 920.0;
  p->WatMHeat = 333000.0;  p->PorEvaF = 1.0;
  p->ZRefW = 10.0;  p->ZRefT = 2.0;  p->ZeroDisp = 0.0;  p->ZMom = 0.4;  p->ZHeat = 0.001;
  p->Emiss = 0.95;  p->Albedo = 0.10;  p->Albedo_surroundings = 0.15;  p->MaxPormms = 1.0;
  p->TClimG = 6.4;  p->DampDpth = 2.7;  p->Omega = 2.0 * 3.14159265358979323846 / 365.0;  p->AZ = 0.6;
  p->DampWearF = 0.5;  p->AlbDry = 0.1;  p->AlbSnow = 0.6;
  p->vsh1 = 1.94e6;  p->vsh2 = 1.28e6;  p->Poro1 = 0.1;  p->Poro2 = 0.4;
  p->RhoB1 = 2.11;  p->RhoB2 = 1.6;  p->Silt1 = 0.1;  p->Silt2 = 0.8;
  p->freezing_limit_normal = -0.25;  p->snow_melting_limit_normal = 0.25;  p->ice_melting_limit_normal = 0.25;
  p->frost_melting_limit_normal = 1.25;  p->frost_formation_limit_normal = 0.25;  p->T4Melt_normal = 0.25;
  p->TLimColdH = -19.0;  p->TLimColdL = -21.0;  p->WetSnowFormR = 0.1;  p->WetSnowMeltR = 0.6;
  p->PLimSnow = 0.3;  p->PLimRain = 0.7;
  p->MaxSnowmms = 100.0;  p->MaxDepmms = 2.0;  p->MaxIcemms = 50.0;  p->MaxExtmms = 1.0;
  p->MissValI = -9999.0;  p->MissValR = -99.99;  p->Snow2IceFac = 0.5;
  // examples/example1/src/InputParameters.cpp:11-22
  p->MinPrecmm = 0.05 * DTSecs / 3600.0;
  p->MinWatmms = 0.01 * DTSecs / 3600.0;
  p->MinSnowmms = 0.1 * DTSecs / 3600.0;
  p->MaxWatmms = p->MaxPormms + p->MaxExtmms;
  p->WDampLim = 0.1 * p->MaxPormms;
  p->WWetLim = 0.9 * p->MaxPormms;
  p->WWearLim = 0.1 * p->MaxPormms;
  p->MinDepmms = 0.01 * DTSecs / 3600.0;
  p->MinIcemms = 0.05 * DTSecs / 3600.0;
}

extern "C" void roadsurf_default_settings(InputSettings* s, int SimLen, double DTSecs)
{
  if (!s) return;
  std::memset(s, 0, sizeof *s);
  // examples/example1/src/InputSettings.h:13-23
  s->SimLen = SimLen;
  s->use_coupling = 0;
  s->use_relaxation = 0;
  s->force_tsurf = 0;
  s->DTSecs = DTSecs;
  s->tsurfOutputDepth = -9999.9;
  s->NLayers = 15;
  s->coupling_minutes = 180;
  s->couplingEffectReduction = 4.0 * 3600.0;
  s->outputStep = 60;
}

int rs_build_model(const InputSettings* s, const InputParameters* p, RsModel* m, char* err, int errlen)
{
  std::memset(m, 0, sizeof *m);
  if (s->NLayers < RS_MIN_LAYERS || s->NLayers > RS_MAX_LAYERS)
  {
    std::snprintf(err, errlen, "NLayers=%d outside supported range [%d, %d]", s->NLayers,
                  RS_MIN_LAYERS, RS_MAX_LAYERS);
    return RS_ERR_UNSUPPORTED;
  }
  if (!(s->DTSecs > 0.0) || s->SimLen < 1)
  {
    std::snprintf(err, errlen, "bad settings: DTSecs=%g SimLen=%d", s->DTSecs, s->SimLen);
    return RS_ERR_BAD_ARGUMENT;
  }
  const int n = s->NLayers;
  m->nlayers = n;
  m->use_coupling = (s->use_coupling == 1);      // int2Logical, src/Initialization.f90:560-571
  m->use_relaxation = (s->use_relaxation == 1);
  m->force_tsurf = (s->force_tsurf == 1);
  m->sim_len_hint = s->SimLen;
  m->DT = s->DTSecs;
  m->Tph = s->DTSecs / 3600.0;
  m->tsurfOutputDepth = s->tsurfOutputDepth;
  m->couplingEffectReduction = s->couplingEffectReduction;
  m->coupling_span_real = static_cast<double>(s->coupling_minutes * 60) / s->DTSecs;
  m->coupling_span = static_cast<int>(m->coupling_span_real);

  m->NightOn = p->NightOn;
  m->NightOff = p->NightOff;
  m->CalmLimDay = p->CalmLimDay;
  m->CalmLimNgt = p->CalmLimNgt;
  m->TrfFricNgt = p->TrfFricNgt;
  m->TrFfricDay = p->TrFfricDay;

  m->Grav = p->Grav;
  m->SB_Const = p->SB_Const;
  m->VK_Const = p->VK_Const;
  m->LVap = p->LVap;
  m->LFus = p->LFus;
  m->ZRefT = p->ZRefT;
  m->Emiss = p->Emiss;
  m->Albedo0 = p->Albedo;
  m->Albedo_surroundings = p->Albedo_surroundings;
  m->MaxPormms = p->MaxPormms;
  m->logMom = std::log((p->ZRefW + p->ZMom) / p->ZMom);
  m->logHeat = std::log((p->ZRefW + p->ZHeat) / p->ZHeat);
  m->logCond = std::log((p->ZRefW - p->ZeroDisp + p->ZHeat) / p->ZHeat);
  m->logUstar = std::log((p->ZRefW - p->ZeroDisp + p->ZMom) / p->ZMom);
  m->TClimG = p->TClimG;
  m->AZ = p->AZ;
  m->Omega = p->Omega;
  m->DampDpth = p->DampDpth;
  m->dry1 = (1.0 - p->Poro1) * p->vsh1;
  m->dry2 = (1.0 - p->Poro2) * p->vsh2;

  m->WatDens = p->WatDens;
  m->WatMHeat = p->WatMHeat;
  m->PorEvaF = p->PorEvaF;
  m->DampWearF = p->DampWearF;
  m->TLimFreeze = p->freezing_limit_normal;
  m->TLimMeltSnow = p->snow_melting_limit_normal;
  m->TLimMeltIce = p->ice_melting_limit_normal;
  m->TLimMeltDep = p->frost_melting_limit_normal;
  m->TLimDew = p->frost_formation_limit_normal;
  m->T4Melt0 = p->T4Melt_normal;
  m->WetSnowFormR = p->WetSnowFormR;
  m->WetSnowMeltR = p->WetSnowMeltR;
  m->PLimSnow = p->PLimSnow;
  m->PLimRain = p->PLimRain;
  m->MinPrecmm = p->MinPrecmm;
  m->MinWatmms = p->MinWatmms;
  m->MinSnowmms = p->MinSnowmms;
  m->MinDepmms = p->MinDepmms;
  m->MinIcemms = p->MinIcemms;
  m->MaxSnowmms = p->MaxSnowmms;
  m->MaxDepmms = p->MaxDepmms;
  m->MaxIcemms = p->MaxIcemms;
  m->MaxWatmms = p->MaxWatmms;
  m->AlbDry = p->AlbDry;
  m->AlbSnow = p->AlbSnow;
  m->MissValI = p->MissValI;
  m->WWetLim = p->WWetLim;
  m->WWearLim = p->WWearLim;

  // layer depths: thickness grows geometrically (single-precision product, see header comment)
  const double zadd = f4(0.02f);
  m->ZDpth[1] = 0.0;
  for (int i = 1; i <= n; ++i)
    m->ZDpth[i + 1] = m->ZDpth[i] + f4(0.0103f * powi_f32(1.4f, i - 1)) + zadd;

  m->DyC[1] = (m->ZDpth[2] - m->ZDpth[1]) / 2.0;
  for (int j = 2; j <= n; ++j) m->DyC[j] = (m->ZDpth[j + 1] - m->ZDpth[j - 1]) / 2.0;
  for (int j = 1; j <= n; ++j)
  {
    m->WCont[j] = (j <= 2) ? f4(0.01f) : f4(0.3f);
    const double cc = (j <= 2) ? campbell_cc(p->RhoB1, p->Silt1, m->WCont[j])
                               : campbell_cc(p->RhoB2, p->Silt2, m->WCont[j]);
    const double dyk = m->ZDpth[j + 1] - m->ZDpth[j];
    m->condDZ[j] = -(cc / dyk);
  }
  m->hs1_dz = m->ZDpth[2] - m->ZDpth[1];
  m->two_dt = 2.0 * s->DTSecs;
  m->inv_two_dt = 1.0 / m->two_dt;
  m->inv_DT = 1.0 / s->DTSecs;
  m->inv_3600 = 1.0 / 3600.0;
  m->inv_1000 = 1.0 / 1000.0;
  m->inv_3364 = 1.0 / 3364.0;
  m->inv_1p5 = 1.0 / 1.5;
  m->inv_4h = 1.0 / static_cast<double>(4.f * 3600.f);
  m->inv_CER = 1.0 / s->couplingEffectReduction;
  m->WatMHeatDens = p->WatMHeat * p->WatDens;
  m->inv_WatMHeatDens = 1.0 / m->WatMHeatDens;

  // fixed output depth, resolved once (getTempAtDepth with a run-constant depth)
  m->depth_mode = 0;
  m->depth_idx = 0;
  const double d = s->tsurfOutputDepth;
  if (d >= 0.0)
  {
    if (std::fabs(d - 0.0) < f4(0.00001f))
      m->depth_mode = 1;
    else if (d > m->ZDpth[n + 1])
      m->depth_mode = 2;
    else
    {
      m->depth_mode = 3;
      int idx = 1;
      for (; idx <= n; ++idx)
        if (d > m->ZDpth[idx] && d <= m->ZDpth[idx + 1]) break;
      if (idx > n) idx = n;
      m->depth_idx = idx;
    }
  }
  if (err && errlen > 0) err[0] = 0;
  return RS_OK;
}
