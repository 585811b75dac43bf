"""SECOND, independently written restatement of the reference's per-point simulation loop -- plain Python,
one point at a time, written from the Fortran sources (not from roadsurf_oracle.hpp) as an N-version
check on the C++ oracle.  TEST INFRASTRUCTURE ONLY (tests/test_second_restatement.py); far too slow for
anything but the golden cases.  PARITY UNPINNED like the oracle itself: no gfortran here.

Conventions: every function names the Fortran routine it follows (file:line of the reference tree).
Un-suffixed Fortran real literals are REAL(4): written r4(x) = float(numpy.float32(x)).  Arrays that are
1-based in Fortran are Python lists with index 0 unused (Tmp, TmpNw, GCond are 0-based there as well).
"""
import math

import numpy as np


def r4(x):
    return float(np.float32(x))


def pow_r4_i4(x, n):
    """REAL(4) ** INTEGER as libgfortran's _gfortran_pow_r4_i4 evaluates it: binary exponentiation in single
    precision (n >= 0 here)."""
    x, result = np.float32(x), np.float32(1.0)
    while True:
        if n & 1:
            result = np.float32(result * x)
        n >>= 1
        if not n:
            return result
        x = np.float32(x * x)


F32 = np.float32
MISS = -9999.0


class Failed(Exception):
    pass


class Point:
    """All derived-type state of one point (src/*.f90.inc), flattened into attributes."""

    def __init__(self, inp, settings, params, local):
        """inp: dict name -> 1-D arrays of length SimLen (Tair, Tdew, VZ, Rhz, prec, SW, LW, SW_dir, LW_net,
        TSurfObs, PrecPhase, Depth, year..second) + local_horizons[360]; arrays are copied (the reference
        mutates its inputs).  settings / params / local: the ctypes interop structs."""
        self.inp = {k: np.array(v, copy=True) for k, v in inp.items()}
        self.S, self.P, self.L = settings, params, local
        n = settings.SimLen
        self.out = {k: np.full(n, MISS) for k in ("TsurfOut", "SnowOut", "WaterOut", "IceOut", "DepositOut", "Ice2Out")}
        self.executed = 0

    # ------------------------------------------------------------------ src/Initialization.f90
    def initialization(self):
        S, P, L, inp = self.S, self.P, self.L, self.inp
        # initSettings :442-476
        self.SimLen, self.InitLenI, self.DT = S.SimLen, L.InitLenI, S.DTSecs
        self.tsurfOutputDepth, self.N = S.tsurfOutputDepth, S.NLayers
        self.use_coupling, self.use_relaxation = S.use_coupling == 1, S.use_relaxation == 1
        self.force_tsurf = S.force_tsurf == 1
        # initOutputArrays :397-412 (done in __init__); setInputParam, src/InputOutput.f90:4-39
        self.TairR, self.VZR, self.RhzR = float(F32(L.tair_relax)), float(F32(L.VZ_relax)), float(F32(L.RH_relax))
        if (self.TairR < r4(-100.0) or self.TairR > r4(100.0) or self.VZR < 0.0 or self.VZR > 100.0 or self.RhzR < 0.0
                or self.RhzR > 110):
            self.use_relaxation = False
        self.obsI1, self.obsTsurf1, self.lastTsurfObs = L.couplingIndexI, L.couplingTsurf, L.couplingTsurf
        if L.couplingTsurf < -100 or self.obsI1 < 1:
            self.use_coupling = False
        # initVariablesAndParameters :66-147
        self.failed = False
        self.Tph = self.DT / 3600.0
        N = self.N
        # initDepth :217-235  (0.0103*1.4**(I-1) is a REAL(4) expression; ZAdd = 0.02 is a REAL(4) literal in a REAL(8))
        Z = [0.0] * (N + 2)
        zadd = r4(0.02)
        for i in range(1, N + 1):
            Z[i + 1] = Z[i] + float(F32(0.0103) * pow_r4_i4(1.4, i - 1)) + zadd
        self.Z = Z
        # initSurf :290-308
        self.Q2Melt, self.EvapmmTS = 0.0, 0.0
        self.Wat = self.Snow = self.Ice = self.Ice2 = self.Dep = 0.0
        # InitParam :310-358
        self.logMom = math.log((P.ZRefW + P.ZMom) / P.ZMom)
        self.logHeat = math.log((P.ZRefW + P.ZHeat) / P.ZHeat)
        self.logCond = math.log((P.ZRefW - P.ZeroDisp + P.ZHeat) / P.ZHeat)
        self.logUstar = math.log((P.ZRefW - P.ZeroDisp + P.ZMom) / P.ZMom)
        self.Albedo = P.Albedo
        # initTemp :238-287
        Tmp = [0.0] * (N + 2)
        Tmp[0] = inp["Tair"][0]
        first = inp["TSurfObs"][0] if inp["TSurfObs"][0] > -100 else inp["Tair"][0]
        for i in range(1, 5):
            Tmp[i] = first
        juld = self.julday(0)
        Tmp[N + 1] = P.TClimG + P.AZ * math.sin(P.Omega * juld + P.Omega * (-170) - (Z[N + 1] / P.DampDpth))
        for i in range(5, N + 1):
            Tmp[i] = Tmp[4] + (Tmp[N + 1] - Tmp[4]) / (Z[N + 1] - Z[4]) * (Z[i] - Z[4])
        self.Tmp, self.TmpNw = Tmp, list(Tmp)
        d0 = inp["Depth"][0]
        self.Ts = self.temp_at_depth(d0) if d0 >= 0 else 0.5 * (Tmp[1] + Tmp[2])
        # initVariables :360-395 (only what is read later)
        self.BLCond = r4(-99.9)
        self.TairInitEnd = self.VZInitEnd = self.RhzInitEnd = r4(-99.9)
        # HCapValues, src/BalanceModel.f90:158-186
        def fc(rhob, silt):
            a = r4(0.65) - r4(0.78) * rhob + r4(0.60) * rhob * rhob
            b = r4(1.06) * rhob
            c = 1 + r4(2.6) / math.sqrt(silt) if silt > r4(0.00001) else 0.0
            d = r4(0.03) + r4(0.1) * rhob * rhob
            return a, b, c, d, 4.0
        f1, f2 = fc(P.RhoB1, P.Silt1), fc(P.RhoB2, P.Silt2)
        # ground_prop_init :181-214
        DyC, DyK, W = [0.0] * (N + 2), [0.0] * (N + 2), [0.0] * (N + 2)
        DyC[1] = (Z[2] - Z[1]) / 2.0
        for j in range(2, N + 1):
            DyC[j] = (Z[j + 1] - Z[j - 1]) / 2.0
        for j in range(1, N + 1):
            DyK[j] = Z[j + 1] - Z[j]
        W[1] = W[2] = r4(0.01)
        for j in range(3, N + 1):
            W[j] = r4(0.3)
        self.DyC, self.DyK, self.W = DyC, DyK, W
        # CalcCC, src/BalanceModel.f90:254-279
        CC = [0.0] * (N + 2)
        for j in range(1, N + 1):
            a, b, c, d, e = f1 if j <= 2 else f2
            CC[j] = a + b * W[j] - (a - d) * math.exp(-(c * W[j]) ** e)
        self.CC = CC
        self.VSH, self.HS = [0.0] * (N + 2), [0.0] * (N + 2)
        self.condDZ, self.capDZ = [0.0] * (N + 2), [0.0] * (N + 2)
        self.calc_hcap()
        self.calc_capdz()
        # initCoupling, src/Coupling.f90:144-169
        self.it = 0
        self.TsNA = self.TsNB = self.RcNA = self.RcNB = -9999.0
        self.RadCoeff = self.RadCoeffPrev = 1.0
        self.SwCof = self.LwCof = 1.0
        self.again, self.cfailed = False, False
        self.SWcorr = self.LWcorr = 0.0
        self.inCpl = False
        self.Tsurf_end_coup1 = 0.0
        # initCouplingTimes :486-534
        self.cstart = self.cend = -99
        if self.use_coupling and self.obsI1 > -1:
            self.cend = self.obsI1
            if self.obsI1 <= S.coupling_minutes * 60 / self.DT:
                self.cstart = 1
            else:
                self.cstart = self.obsI1 - int(S.coupling_minutes * 60 / self.DT)
        else:
            self.use_coupling = False
        # condInit :479-557
        self.T4Melt = P.T4Melt_normal
        self.Snow2IceFac = P.Snow2IceFac
        if inp["VZ"][0] < r4(0.4):
            inp["VZ"][0] = r4(0.4)
        self.Tair, self.VZ, self.Rhz = inp["Tair"][0], inp["VZ"][0], inp["Rhz"][0]
        self.Ts = self.temp_at_depth(d0) if d0 >= 0 else (Tmp[1] + Tmp[2]) / 2.0
        self.boundary_layer()
        if self.lastTsurfObs < -100:
            self.cfailed = True

    # ------------------------------------------------------------------ src/BalanceModel.f90
    def julday(self, k):  # :325-351
        mon_end = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 0, 31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335]
        y, m, d = int(self.inp["year"][k]), int(self.inp["month"][k]), int(self.inp["day"][k])
        leap = 1 - min(y % 4, 1) + min(y % 100, 1) - min(y % 400, 1)
        return mon_end[m + leap * 12 - 1] + d

    def temp_at_depth(self, depth):  # :390-417 (reads ground%Tmp)
        Z, T, zlen = self.Z, self.Tmp, self.N + 1
        if abs(depth - 0.0) < r4(0.00001):
            return T[1]
        if depth > Z[zlen]:
            return T[zlen]
        idx = zlen
        for k in range(1, zlen):
            if depth > Z[k] and depth <= Z[k + 1]:
                idx = k
                break
        return T[idx] + (depth - Z[idx]) * (T[idx + 1] - T[idx]) / (Z[idx + 1] - Z[idx])

    def calc_hcap(self):  # CalcHCapHCond :189-251
        P, N = self.P, self.N
        for i in range(1, N + 1):
            t = self.TmpNw[i]
            if t >= 0:
                t2 = t * t
                roo = -r4(0.0050) * t2 + r4(0.0079) * t + r4(1000.0028)
                cw = r4(0.0000102) * t2 * t2 - r4(0.0017169) * t2 * t + r4(0.11516) * t2 - r4(3.4739) * t + r4(4217.2)
            else:
                roo, cw = r4(920.0), r4(2100.0)
            chwt = roo * cw
            if i <= 2:
                self.VSH[i] = (1.0 - P.Poro1) * P.vsh1 + self.W[i] * chwt
            else:
                self.VSH[i] = (1.0 - P.Poro2) * P.vsh2 + self.W[i] * chwt
            lo = self.Z[i] if i == 1 else self.Z[i - 1]
            self.HS[i] = self.VSH[i] * (self.Z[i + 1] - lo) / (2.0 * self.DT)

    def calc_capdz(self):  # calcCapDZCondDZ :132-155
        for j in range(1, self.N + 1):
            self.condDZ[j] = -(self.CC[j] / self.DyK[j])
            self.capDZ[j] = -(1 / (self.DyC[j] * self.VSH[j]))

    def balance_one_step(self, i):  # BalanceModelOneStep :7-86 (i = 1-based input index)
        S, P, inp, N = self.S, self.P, self.inp, self.N
        k = i - 1
        # SetDayDependendVariables :354-387
        hour = int(inp["hour"][k])
        if hour >= P.NightOn or hour <= P.NightOff:
            calm, self.TrfFric = P.CalmLimNgt, P.TrfFricNgt
        else:
            calm, self.TrfFric = P.CalmLimDay, P.TrFfricDay
        if self.VZ < calm:
            self.VZ = calm
        self.boundary_layer()
        # CalcRNet :282-307
        tk = self.Ts + r4(273.15)
        tk2 = tk * tk
        rbb = P.Emiss * P.SB_Const * (tk2 * tk2)
        rnet = (1. - self.Albedo) * inp["SW"][k] * self.SwCof + P.Emiss * inp["LW"][k] * self.LwCof - rbb
        self.calc_hcap()
        self.calc_capdz()
        # calcProfile :90-129
        T = self.Tmp
        G = [0.0] * (N + 2)
        G[0] = rnet - self.LE + self.TrfFric + self.BLCond * (T[0] - T[1])
        self.TmpNw = list(T)
        for j in range(1, N + 1):
            G[j] = self.condDZ[j] * (T[j + 1] - T[j])
        for j in range(1, N + 1):
            self.TmpNw[j] = T[j] + self.DT * (self.capDZ[j] * (G[j] - G[j - 1]))
        # calcHStor :311-322
        t1 = (T[1] + 3. * T[2]) / 4.
        tn1 = (self.TmpNw[1] + 3. * self.TmpNw[2]) / 4.
        self.HStor = self.HS[1] * (tn1 - t1)
        depth = self.tsurfOutputDepth if self.tsurfOutputDepth >= 0.0 else inp["Depth"][k]
        self.melting()
        self.Tmp = list(self.TmpNw)
        self.Ts = self.temp_at_depth(depth) if depth >= 0 else (self.Tmp[1] + self.Tmp[2]) / 2.0

    # ------------------------------------------------------------------ src/BoundaryLayer.f90
    def boundary_layer(self):  # CalcBLCondAndLE :3-109, calcRaero :112-131, CalcLE :134-190
        P = self.P
        Ts, Tair, VZ, Rhz = self.Ts, self.Tair, self.VZ, self.Rhz
        conv = r4(0.001)
        blc = self.BLCond
        tak = Tair + r4(273.15)
        dens = 100000.0 / (r4(287.05) * tak)
        hcap = 1005.0 + ((tak - 250.0) * (tak - 250.0)) / 3364.   # (x)**2 is x*x in Fortran
        vcap = hcap * dens
        psy = r4(0.1) * (r4(0.00063) * tak + r4(0.47496))
        watden = -r4(0.0050) * Ts * Ts + r4(0.0079) * Ts + r4(1000.0028)
        psim = psih = 0.0
        old = blc
        j = 0
        for j in range(1, 41):
            old = blc
            ustar = P.VK_Const * VZ / (self.logUstar + psim)
            blc = vcap * P.VK_Const * ustar / (self.logCond + psih)
            stab = -P.VK_Const * P.ZRefT * P.Grav * blc * (Ts - Tair) / (vcap * (Tair + r4(273.15)) * (ustar * ustar * ustar))
            if stab > 1:
                stab = 1
            if stab > 0:
                psih = r4(4.7) * stab
                psim = psih
            else:
                psih = -2.0 * math.log((1.0 + math.sqrt(1.0 - 16.0 * stab)) / 2.0)
                psim = r4(0.6) * psih
            if abs(blc - old) < conv and j >= 5:
                break
        raero = (self.logMom + psim) * (self.logHeat + psih) / (P.VK_Const * P.VK_Const * VZ)
        if raero > 30.0:
            raero = 30.
        def esat(t):
            if t < 0:
                return r4(0.61078) * math.exp(r4(21.875) * t / (t + r4(265.5)))
            return r4(0.61078) * math.exp(r4(17.269) * t / (t + r4(237.3)))
        esurf = esat(Ts)
        eair = min(r4(0.01) * Rhz, 1.0) * esat(Tair)
        le = (dens * hcap * (esurf - eair)) / (psy * raero)
        if Ts >= 0.0:
            evap = (le / (P.LVap * watden)) * 1000.0 * self.DT
        else:
            evap = (le / (P.LFus * watden)) * 1000.0 * self.DT
        if le > 0.0 and self.Wat <= 0.0:
            le, evap = 0.0, 0.0
        self.LE, self.EvapmmTS, self.BLCond = le, evap, blc

    # ------------------------------------------------------------------ src/Storage.f90
    def melting(self):  # :319-402 (its TsurfAve result is overwritten by BalanceModelOneStep)
        if self.Snow > 0.0 or self.Ice > 0.0 or self.Ice2 > 0.0:
            if (self.HStor <= r4(0.00001) or self.Ts <= self.T4Melt or self.Q2Melt <= 0
                    or (self.inCpl and self.lastTsurfObs < self.T4Melt)):
                if self.Ts < 0.5:
                    self.Q2Melt = 0.0
                    return
                elif self.Ts > 2.0:
                    q = self.HS[1] * (self.TmpNw[1] - self.T4Melt)
                    if q < self.Q2Melt:
                        self.Q2Melt = q
                    return
            q = self.HS[1] * (self.TmpNw[1] - self.T4Melt)
            if self.Q2Melt >= q:
                self.Q2Melt = q
                self.TmpNw[1] = self.T4Melt + r4(0.01)
                self.TmpNw[2] = self.T4Melt + r4(0.01)
            else:
                left = q - self.Q2Melt
                self.TmpNw[1] = self.T4Melt + (left / self.HS[1])
                self.TmpNw[2] = self.T4Melt + r4(0.01)
        else:
            self.Q2Melt = 0.0

    def precipitation_to_storage(self, i):  # :9-29 + CalcPrecType, src/Cond.f90:143-249
        P = self.P
        phase = int(self.inp["PrecPhase"][i - 1])
        rain = snow = 0.0
        interp = True
        if phase > P.MissValI:
            interp = False
            if self.PrecInTStep <= P.MinPrecmm:
                self.PrecInTStep = 0.0
            elif phase in (0, 1, 4, 5):          # none, rain, freezing drizzle, freezing rain (src/Constants.h)
                rain = self.PrecInTStep
            elif phase == 2:                     # sleet
                snow = self.PrecInTStep / 2.
                rain = snow
            elif phase in (3, 6):                # snow, hail
                snow = self.PrecInTStep
            else:
                interp = True
        if interp:
            if self.PrecInTStep <= P.MinPrecmm:
                self.PrecInTStep = 0.0
                rain = snow = 0.0
            else:
                snow = 0.0
                pexp = 22.0 - r4(2.7) * self.Tair - r4(0.20) * self.Rhz
                prain = 1.0 / (1.0 + math.exp(pexp))
                if prain < P.PLimSnow:
                    snow = self.PrecInTStep
                elif prain > P.PLimRain:
                    rain = self.PrecInTStep
                else:
                    snow = self.PrecInTStep / 2.
                    rain = snow
        self.Wat = self.Wat + rain
        self.Snow = self.Snow + snow

    def road_cond(self):  # WearFactors src/Cond.f90:69-103, RoadCond :9-65, storages src/Storage.f90:33-314, :409-432
        P, DT, Tph = self.P, self.DT, self.Tph
        MaxPor = P.MaxPormms
        snowtran = float(F32(0.2) + F32(0.25)) * self.Snow
        snowtran = max(snowtran, r4(0.01))
        if self.Snow < r4(0.2):
            snowtran = snowtran * 3
        self.Snow2IceFac = float(F32(0.25) / (F32(0.2) + F32(0.25)))
        snowtran = snowtran * Tph
        icewear = max(float(F32(1.1) * F32(2.0) * F32(0.145)) * self.Ice, r4(0.01)) * Tph
        icewear2 = max(float(F32(1.1) * F32(2.0) * (F32(4.0) * F32(0.290))) * self.Ice2, r4(0.01)) * Tph
        depwear = max(float(F32(0.5) * F32(2.0) * (F32(4.0) * F32(0.290))) * self.Dep, r4(0.01)) * Tph
        watwear = 10 * max(r4(0.145) * self.Wat, r4(0.06)) * Tph
        # WaterStorage
        if self.Snow <= 0.0 and self.Ice <= 0.0 and self.Dep <= 0.0 and self.Ts > P.frost_formation_limit_normal:
            if self.Wat > MaxPor:
                self.Wat = self.Wat - self.EvapmmTS
            else:
                self.Wat = self.Wat - P.PorEvaF * self.EvapmmTS
        if self.Wat > 0.0:
            if self.Wat < P.WWearLim:
                watwear = 0.0
            if self.Wat > P.WWetLim:
                self.Wat = self.Wat - watwear
            else:
                self.Wat = self.Wat - P.DampWearF * watwear
        if self.Wat < P.MinWatmms:
            self.Wat = 0.0
        if self.Wat > P.MaxWatmms:
            self.Wat = P.MaxWatmms
        ext = max(self.Wat - MaxPor, 0.)
        # SnowStorage
        tot = ext + self.Snow
        ratio = ext / tot if tot > r4(0.001) else 0.0
        wet = False
        if self.Snow > 0.0:
            if ratio > P.WetSnowFormR:
                wet = True
            if self.Dep > 0.0:
                self.Ice = self.Ice + self.Dep
                self.Dep = 0.0
            if self.Q2Melt > 0.0 and self.Ts >= P.snow_melting_limit_normal:
                melted = (self.Q2Melt * DT) / (P.WatMHeat * P.WatDens)
                self.Snow = self.Snow - 1000. * melted
                self.Wat = self.Wat + 1000. * melted
        if self.Snow > 0.0:
            self.Snow = self.Snow - snowtran
            self.Ice = self.Ice + self.Snow2IceFac * snowtran
            self.Ice2 = self.Ice2 + self.Snow2IceFac * snowtran
        if self.Snow > 0.0 and wet:
            if ratio > P.WetSnowMeltR:
                self.Wat = self.Wat + self.Snow
                self.Snow = 0.0
                wet = False
            if self.Ts < P.freezing_limit_normal:
                self.Ice = self.Ice + self.Snow + self.Wat
                self.Ice2 = self.Ice2 + self.Snow + self.Wat
                self.Snow = 0.0
                self.Wat = 0.0
        if self.Snow < P.MinSnowmms:
            self.Snow = 0.0
        if self.Snow > P.MaxSnowmms:
            self.Snow = self.Snow - (P.MaxSnowmms / 2.)
        # IceStorage
        if self.Ts < P.freezing_limit_normal and self.Wat > 0.0:
            self.Ice = self.Ice + self.Wat
            self.Ice2 = self.Ice2 + self.Wat
            self.Wat = 0.0
        if self.Snow <= 0. and self.Ice > 0.:
            if self.Q2Melt > 0.0 and self.Ts >= P.ice_melting_limit_normal:
                melted = (self.Q2Melt * DT) / (P.WatMHeat * P.WatDens)
                self.Ice = self.Ice - 1000. * melted
                self.Ice2 = self.Ice2 - 1000. * melted
                self.Wat = self.Wat + 1000. * melted
        if self.Ice > 0.:
            self.Ice = self.Ice - icewear
        if self.Ice2 > 0.:
            self.Ice2 = self.Ice2 - icewear2
        if self.Ice < P.MinIcemms:
            self.Ice = 0.0
        if self.Ice > P.MaxIcemms:
            self.Ice = P.MaxIcemms
        if self.Ice2 < P.MinIcemms:
            self.Ice2 = 0.0
        if self.Ice2 > P.MaxIcemms:
            self.Ice2 = P.MaxIcemms
        # DepositStorage
        if self.EvapmmTS < 0.0:
            self.Dep = self.Dep - self.EvapmmTS
        if self.Ts > P.frost_melting_limit_normal:
            self.Wat = self.Wat + self.Dep
            self.Dep = 0.0
        if self.Snow <= 0.0 and self.Dep > 0:
            self.Dep = self.Dep - depwear
        if self.Dep < P.MinDepmms:
            self.Dep = 0.0
        if self.Dep > P.MaxDepmms:
            self.Wat = self.Wat + (self.Dep - P.MaxDepmms)
            self.Dep = P.MaxDepmms
        if self.Wat < P.MinWatmms:
            self.Wat = 0.0
        if self.Wat > P.MaxWatmms:
            self.Wat = P.MaxWatmms
        # NewMeltFreezeHeat
        self.Q2Melt = 0.0
        if self.Snow > 0.0:
            self.Q2Melt = P.WatMHeat * P.WatDens * (self.Snow / 1000.) / DT
            self.T4Melt = P.snow_melting_limit_normal
        if self.Snow <= 0.0 and self.Ice > 0.0:
            self.Q2Melt = P.WatMHeat * P.WatDens * (self.Ice / 1000.) / DT
            self.T4Melt = P.ice_melting_limit_normal
        if self.Q2Melt < 0.0:
            self.Q2Melt = 0.0
        # CalcAlbedo, src/Cond.f90:105-139
        icesum = max(0.5 * (self.Ice + self.Ice2) + self.Dep, 0.0)
        self.Albedo = P.AlbDry
        if self.Snow > r4(0.01) and self.Snow > self.Ice:
            self.Albedo = P.AlbSnow
        elif self.Ice > r4(0.01) or self.Dep > r4(0.01):
            self.Albedo = P.AlbDry + (icesum / 1.5) * (P.AlbSnow - P.AlbDry) if icesum < 1.5 else P.AlbSnow

    # ------------------------------------------------------------------ src/SunPosition.f90, src/ModRadiation.f90
    def sun_position(self, i):  # SunPosition :4-18, JulianEphemerisDay :196-260, calcElevationAzimuth :20-194
        inp, k = self.inp, i - 1
        yr_i, mon_i = int(inp["year"][k]), int(inp["month"][k])
        if mon_i <= 2:
            yr, mo = float(F32(yr_i - 1)), float(F32(mon_i + 12))
        else:
            yr, mo = float(F32(yr_i)), float(F32(mon_i))
        day = float(F32(int(inp["day"][k])) + F32(int(inp["hour"][k])) / F32(24.)
                    + F32(int(inp["minute"][k])) / (F32(24.) * F32(60.))
                    + F32(int(inp["second"][k])) / (F32(24.) * F32(60.) * F32(60.)))
        A = math.trunc(yr / 100.)
        B = 2. - A + math.trunc(A / 4.)
        jde = math.trunc(365.25 * (yr + 4716)) + math.trunc(r4(30.6001) * (mo + 1.)) + day + B - 1.5245e3
        pi = 4 * math.atan(1.0)
        T = (jde - 2451545.0) / (365.25 * 100.)
        def wrap(x, period=360.):
            if x < 0.:
                x = x - period * (math.trunc(x / period) - 1.)
            if x > period:
                x = x - period * math.trunc(x / period)
            return x
        ml = wrap(r4(280.46645) + r4(36000.76983) * T + r4(0.0003032) * T * T)
        ma = wrap(r4(357.52910) + r4(35999.05030) * T - r4(0.0001559) * T * T - r4(0.00000048) * T * T * T)
        sunc = ((r4(1.913600) - r4(0.004817) * T - r4(0.000014) * T * T) * math.sin(ma * pi / 180.)
                + (r4(0.019993) - r4(0.000101) * T) * math.sin(2. * ma * pi / 180.) + r4(0.000290) * math.sin(3. * ma * pi / 180.))
        al = ml + sunc - r4(0.00569) - r4(0.00478) * math.sin((r4(125.04) - r4(1934.136) * T) * pi / 180.)
        al = al * pi / 180.
        tilt = r4(23.43929111) - r4(0.013004166) * T - r4(0.001638888) * T * T + r4(0.005036111) * T * T * T
        eps = (tilt + r4(0.00256) * math.cos((r4(125.04) - r4(1934.136) * T) * pi / 180.)) * pi / 180.
        ra = wrap(math.atan2(math.cos(eps) * math.sin(al), math.cos(al)), 2. * pi)
        decl = math.asin(math.sin(eps) * math.sin(al))
        stg = wrap(r4(280.46061837) + r4(360.98564736629) * (jde - 2451545.0) + r4(0.000387933) * T * T - T * T * T / 38710000.)
        stg = stg * pi / 180.
        cd, sd = math.cos(decl), math.sin(decl)
        latr = pi * self.L.lat / 180.
        sl, cl = math.sin(latr), math.cos(latr)
        ha = stg + self.L.lon * pi / 180. - ra
        cosah = math.cos(ha)
        ce = sd * sl + cd * cl * cosah
        if 1.0 <= ce < r4(1.001):
            chi = 0.
        elif ce >= r4(1.001):
            raise Failed("stop")
        elif r4(-1.001) < ce <= -1.0:
            chi = pi
        else:
            chi = math.acos(ce)
        elev = 90.0 - chi * (180. / pi)
        if ha < 0.:
            ha = 2 * pi + ha
        elif ha > 2 * pi:
            ha = ha - 2 * pi
        if elev > 0:
            cosele = math.cos((pi / 2.0) - chi)
            if r4(-0.0001) <= cosele < r4(0.0001):
                az = r4(-9999.9)
            else:
                pre = (sd * cl - cd * sl * cosah) / cosele
                if 1.0 <= pre < r4(1.001):
                    az = 0.0
                elif pre >= r4(1.001):
                    raise Failed("stop")
                elif r4(-1.001) < pre <= -1.0:
                    az = pi
                else:
                    az = math.acos(pre)
            if ha < pi:
                az = 2 * pi - az
            return elev, az * (180. / pi)
        return r4(-9999.9), r4(-9999.9)

    def mod_radiation(self, i):  # src/ModRadiation.f90:7-73 (rewrites the input arrays)
        inp, k, sv = self.inp, i - 1, self.L.sky_view
        dif = inp["SW"][k] - inp["SW_dir"][k]
        lw_sur = inp["LW_net"][k] - inp["LW"][k]
        elev, az = self.sun_position(i)
        idx = int(round(az)) if az >= 0 else -int(round(-az))        # NINT
        if idx == 360:
            idx = 0
        horizon = inp["local_horizons"][idx] if 0 <= idx < 360 else 0.0   # (out of bounds in the reference when the sun is down)
        shadow = 0.0 if horizon > elev else 1.0
        if elev > 0.0:
            inp["SW_dir"][k] = inp["SW_dir"][k] * shadow
            ref = self.P.Albedo_surroundings * inp["SW_dir"][k] + self.P.Albedo_surroundings * dif
            dif = sv * dif + (1.0 - sv) * ref
            inp["SW"][k] = dif + inp["SW_dir"][k]
        inp["LW"][k] = sv * inp["LW"][k] + (1.0 - sv) * (-lw_sur)

    # ------------------------------------------------------------------ src/InputOutput.f90, Relaxation, Coupling
    def sky_active(self):
        return self.L.sky_view < 1.0 and self.L.sky_view > r4(-0.01)

    def check_values(self, i):  # CheckValues :45-84
        inp, k = self.inp, i - 1
        g = lambda n: inp[n][k]
        if (g("Tair") < -90.0 or g("Tair") > 100.0 or g("Tdew") < -90 or g("Tdew") > 100.0 or g("Rhz") < r4(-0.1)
                or g("Rhz") > 120.0 or g("VZ") < -1.0 or g("VZ") > 100.0 or g("SW") < r4(-0.1) or g("SW") > 4000.0
                or g("LW") < r4(-0.1) or g("LW") > 1000.0 or g("prec") < r4(-0.1) or g("prec") > 500.0):
            self.failed = True
        if self.sky_active():
            if g("SW_dir") < r4(-0.1) or g("SW_dir") > 4000.0 or g("LW_net") < -1000.0 or g("LW_net") > 1000.0:
                self.failed = True
        if inp["SW_dir"][k] > inp["SW"][k]:
            inp["SW_dir"][k] = inp["SW"][k]
        if self.Ts < -100.0 or self.Ts > 100.0:
            self.failed = True

    def set_current_values(self, i):  # SetCurrentValues :86-149
        inp, k = self.inp, i - 1
        self.Tair, self.VZ, self.Rhz = inp["Tair"][k], inp["VZ"][k], inp["Rhz"][k]
        self.PrecInTStep = inp["prec"][k] / 3600 * self.DT
        self.Tmp[0] = self.Tair
        if i <= self.InitLenI or self.force_tsurf:
            if inp["TSurfObs"][k] > -100.0:
                if (not self.use_coupling) or i < self.cstart:
                    self.Tmp[1] = self.Tmp[2] = inp["TSurfObs"][k]
                    depth = self.tsurfOutputDepth if self.tsurfOutputDepth >= 0.0 else inp["Depth"][k]
                    self.Ts = self.temp_at_depth(depth) if depth >= 0 else (self.Tmp[1] + self.Tmp[2]) / 2.0

    def last_values(self):  # lastValues :169-198
        inp, k = self.inp, self.SimLen - 1
        self.Tair, self.VZ, self.Rhz = inp["Tair"][k], inp["VZ"][k], inp["Rhz"][k]
        self.PrecInTStep = inp["prec"][k] / 3600 * self.DT
        self.Tmp[0] = self.Tair
        depth = inp["Depth"][k]
        self.Ts = self.temp_at_depth(depth) if depth >= 0 else (self.Tmp[1] + self.Tmp[2]) / 2.0

    def relaxation(self, i):  # src/Relaxation.f90:10-47
        DT, li = self.DT, self.InitLenI
        if i == li:
            self.TairInitEnd, self.VZInitEnd, self.RhzInitEnd = self.Tair, self.VZ, self.Rhz
        if i > li:
            e = math.exp(-((DT * i) - (DT * li)) / float(F32(4.) * F32(3600.)))
            self.Tair = self.Tair - (self.TairR - self.TairInitEnd) * e
            self.Tmp[0] = self.Tair
            self.VZ = self.VZ - (self.VZR - self.VZInitEnd) * e
            self.Rhz = self.Rhz - (self.RhzR - self.RhzInitEnd) * e
            if self.Rhz > 100.:
                self.Rhz = 100.0

    def coupling_operations1(self, i):  # src/Coupling.f90:10-96; returns the (possibly rewound) i
        inp, P = self.inp, self.P
        self.inCpl = self.cstart <= i <= self.cend
        if i == self.cstart and self.it == 0:
            # saveDataForCoupling :172-210 (SrfIcemms is not saved; Ice2 twice)
            lo, hi = self.cstart - 1, self.cend
            self.save = dict(i=i, Ts=self.Ts, Wat=self.Wat, Ice2=self.Ice2, Dep=self.Dep, Snow=self.Snow, Alb=self.Albedo,
                             Tmp=list(self.Tmp), SW=inp["SW"][lo:hi].copy(), SWd=inp["SW_dir"][lo:hi].copy(),
                             LW=inp["LW"][lo:hi].copy())
            self.SwCof = self.LwCof = 1.0
            self.SWcorr = self.LWcorr = 0.0
        if self.again:
            # uploadDataForCoupling :213-255
            s = self.save
            i = s["i"]
            self.Ts, self.Wat, self.Ice2, self.Dep, self.Snow, self.Albedo = s["Ts"], s["Wat"], s["Ice2"], s["Dep"], s["Snow"], s["Alb"]
            self.Tmp = list(s["Tmp"])
            lo, hi = self.cstart - 1, self.cend
            inp["SW"][lo:hi], inp["SW_dir"][lo:hi], inp["LW"][lo:hi] = s["SW"], s["SWd"], s["LW"]
            self.again = False
            if inp["SW"][i - 1] > inp["LW"][i - 1] and not self.sky_active():
                self.SwCof, self.LwCof = self.RadCoeff, 1.0
            else:
                self.SwCof, self.LwCof = 1.0, self.RadCoeff
        if i > self.cend:
            e = math.exp(-((self.DT * i) - (self.DT * self.cend)) / self.S.couplingEffectReduction)
            self.SwCof = 1.0 + self.SWcorr * e
            self.LwCof = 1.0 + self.LWcorr * e
        if self.inCpl:
            # snowIceCheck :259-289
            o = self.lastTsurfObs
            if o > P.snow_melting_limit_normal and self.Snow > 0.00:
                self.Wat, self.Snow = self.Wat + self.Snow, 0.00
            if o > P.ice_melting_limit_normal and self.Ice > 0.00:
                self.Wat, self.Ice = self.Wat + self.Ice, 0.00
            if o > P.ice_melting_limit_normal and self.Ice2 > 0.00:
                self.Ice2 = 0.00
            if o > P.frost_melting_limit_normal and self.Dep > 0.00:
                self.Wat, self.Dep = self.Wat + self.Dep, 0.00
        return i

    def check_end_coupling(self, i):  # CheckEndCoupling :98-118, CouplingOperations2 :121-141, Coupling_control :292-481
        if not (self.use_coupling and i == self.cend and not self.cfailed):
            return
        if self.it == 0:
            self.Tsurf_end_coup1 = self.Ts
        self.again = False
        K = r4(273.16)
        Ts = self.Ts + K
        obs = self.lastTsurfObs + K
        if not self.cfailed:
            if self.it == 0:
                self.Tsurf_end_coup1 = Ts
            def reset():
                self.SwCof = self.LwCof = 1.0
                self.SWcorr = self.LWcorr = 0.0
            def secant():
                da, db = self.TsNA - obs, obs - self.TsNB
                return self.RcNA - da / (da + db) * (self.RcNA - self.RcNB)
            if self.it == 25:
                if abs(self.Tsurf_end_coup1 - obs) < abs(Ts - obs):
                    self.again = True
                reset()
                self.RadCoeff, self.cfailed = 1.0, True
            elif obs < -100:
                reset()
                self.RadCoeff, self.cfailed, self.again = 1.0, True, True
            elif Ts < 170.0 or Ts > 400.0:
                reset()
                self.cfailed, self.again, self.RadCoeff = True, True, 1.0
            elif Ts - obs > r4(0.1):
                if self.TsNA < -100 or self.TsNA - obs > Ts - obs:
                    self.TsNA, self.RcNA = Ts, self.RadCoeff
                self.again = True
                self.RadCoeff = secant() if (self.TsNA > -100 and self.TsNB > -100) else 0.5 * self.RadCoeff
                if abs(self.RadCoeff - self.RadCoeffPrev) < r4(0.00005):
                    self.TsNA = self.TsNB = -9999
                if self.RadCoeff < r4(0.01):
                    self.RadCoeff, self.cfailed = 1.0, True
                    reset()
                self.RadCoeffPrev = self.RadCoeff
            elif obs - Ts > r4(0.1):
                if self.TsNB < -100 or self.TsNB - obs < Ts - obs:
                    self.TsNB, self.RcNB = Ts, self.RadCoeff
                self.again = True
                self.RadCoeff = secant() if (self.TsNA > -100 and self.TsNB > -100) else 2.0 * self.RadCoeff
                if abs(self.RadCoeff - self.RadCoeffPrev) < r4(0.00005):
                    self.TsNA = self.TsNB = -9999
                self.RadCoeffPrev = self.RadCoeff
            else:
                if self.RadCoeff > 3.0:
                    self.RadCoeff = 1.0
                    reset()
                self.SWcorr, self.LWcorr = self.SwCof - 1.0, self.LwCof - 1.0
                self.cfailed, self.it = False, -1
                self.TsNA = self.TsNB = self.RcNA = self.RcNB = -9999.0
                self.RadCoeff = self.RadCoeffPrev = 1.0
        self.Ts = Ts - K
        self.lastTsurfObs = obs - K
        self.it += 1

    # ------------------------------------------------------------------ examples/example1/src/Simulation.f90
    def one_step(self, i):  # roadModelOneStep :120-172
        self.precipitation_to_storage(i)
        if self.sky_active():
            self.mod_radiation(i)
        self.balance_one_step(i)
        self.road_cond()
        self.executed += 1

    def save_output(self, i):  # SaveOutput, src/InputOutput.f90:151-165
        o, k = self.out, i - 1
        o["SnowOut"][k], o["WaterOut"][k], o["IceOut"][k] = self.Snow, self.Wat, self.Ice
        o["Ice2Out"][k], o["DepositOut"][k], o["TsurfOut"][k] = self.Ice2, self.Dep, self.Ts

    def run(self):  # runsimulation :4-117
        self.initialization()
        i = 1
        while i < self.SimLen and not self.failed:
            self.check_values(i)
            if self.use_coupling:
                i = self.coupling_operations1(i)
            self.set_current_values(i)
            if self.use_relaxation:
                self.relaxation(i)
            self.one_step(i)
            self.save_output(i)
            self.check_end_coupling(i)
            i += 1
        if not self.failed:
            self.last_values()
            self.one_step(self.SimLen)
            self.save_output(i)
        return self.out


def run_point(arrays, settings, params, p):
    """Run point p of a roadsurf_b200.abi.PointArrays; returns (outputs dict, executed steps, failed)."""
    names = {"Tair": "tair", "Tdew": "tdew", "VZ": "VZ", "Rhz": "Rhz", "prec": "prec", "SW": "SW", "LW": "LW",
             "SW_dir": "SW_dir", "LW_net": "LW_net", "TSurfObs": "TSurfObs", "PrecPhase": "PrecPhase", "Depth": "Depth"}
    inp = {k: getattr(arrays, v)[p] for k, v in names.items()}
    t = arrays.time if arrays.time_per_point is None else arrays.time_per_point[p]
    for k, n in enumerate(("year", "month", "day", "hour", "minute", "second")):
        inp[n] = t[k]
    inp["local_horizons"] = arrays.local_horizons[p]
    pt = Point(inp, settings, params, arrays.local[p])
    out = pt.run()
    return out, pt.executed, pt.failed
