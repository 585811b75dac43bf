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
