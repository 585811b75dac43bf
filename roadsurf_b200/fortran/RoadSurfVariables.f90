!> Module RoadSurfVariables for the B200 library: the type names of the reference's module of the
!! same name (src/RoadSurfVariables.f90:13-28), so that a Fortran main program written against the
!! reference -- examples/example1/src/Simulation.f90 is the model case -- compiles unchanged.
!!
!! NOT COMPILED OR TESTED IN THIS REPOSITORY: no Fortran compiler exists in the build image
!! (DESIGN.md section 1).  What the C side of every binding does is tested through the C ABI
!! (tests/test_gpu_parity.py::test_stepwise_*).
!!
!! * The five interoperable types are the ABI (include/roadsurf_b200.h part 1): component order and
!!   kinds are those of src/InputPointers.f90.inc, OutputPointers.f90.inc, InputSettings.f90.inc,
!!   InputParameters.f90.inc and LocalParameters.f90.inc, by necessity.
!! * InputArrays / OutputArrays are the Fortran views of the caller's C arrays
!!   (src/InputArrays.f90.inc, OutputArrays.f90.inc); a main indexes them directly
!!   (Simulation.f90:151,158-159).
!! * The nine state types do NOT mirror the reference's components: the state of a run lives on the
!!   GPU (SoA planes, roadsurf_b200.h RS_STATE_NPLANES).  They carry the session handle plus the few
!!   components an unchanged main reads between calls (Simulation.f90:58,61,79,164-169).
module RoadSurfVariables
   use, intrinsic :: ISO_C_BINDING
   implicit none

   ! ---- ABI types (bind(C)) ------------------------------------------------------------------
   type, bind(C) :: InputPointers
      integer(C_INT) :: inputLen
      type(C_PTR) :: c_tair, c_tdew, c_VZ, c_Rhz, c_prec, c_SW, c_LW, c_SW_dir, c_LW_net, c_TSurfObs
      type(C_PTR) :: c_PrecPhase, c_local_horizons, c_Depth
      type(C_PTR) :: c_year, c_month, c_day, c_hour, c_minute, c_second
   end type InputPointers

   type, bind(C) :: OutputPointers
      integer(C_INT) :: outputLen
      type(C_PTR) :: c_TsurfOut, c_SnowOut, c_WaterOut, c_IceOut, c_DepositOut, c_Ice2Out
   end type OutputPointers

   type, bind(C) :: InputSettings
      integer(C_INT) :: SimLen, use_coupling, use_relaxation, force_tsurf
      real(C_DOUBLE) :: DTSecs, tsurfOutputDepth
      integer(C_INT) :: NLayers, coupling_minutes
      real(C_DOUBLE) :: couplingEffectReduction
      integer(C_INT) :: outputStep
   end type InputSettings

   type, bind(C) :: InputParameters
      real(C_DOUBLE) :: NightOn, NightOff, CalmLimDay, CalmLimNgt, TrfFricNgt, TrFfricDay
      real(C_DOUBLE) :: Grav, SB_Const, VK_Const, LVap, LFus
      real(C_DOUBLE) :: WatDens, SnowDens, IceDens, DepDens, WatMHeat, PorEvaF
      real(C_DOUBLE) :: ZRefW, ZRefT, ZeroDisp, ZMom, ZHeat, Emiss, Albedo, Albedo_surroundings
      real(C_DOUBLE) :: MaxPormms, TClimG, DampDpth, Omega, AZ, DampWearF, AlbDry, AlbSnow
      real(C_DOUBLE) :: vsh1, vsh2, Poro1, Poro2, RhoB1, RhoB2, Silt1, Silt2
      real(C_DOUBLE) :: freezing_limit_normal, snow_melting_limit_normal, ice_melting_limit_normal
      real(C_DOUBLE) :: frost_melting_limit_normal, frost_formation_limit_normal, T4Melt_normal
      real(C_DOUBLE) :: TLimColdH, TLimColdL, WetSnowFormR, WetSnowMeltR, PLimSnow, PLimRain
      real(C_DOUBLE) :: MaxSnowmms, MaxDepmms, MaxIcemms, MaxExtmms, MissValI, MissValR, Snow2IceFac
      real(C_DOUBLE) :: MinPrecmm, MinWatmms, MinSnowmms, MaxWatmms, WDampLim, WWetLim, WWearLim
      real(C_DOUBLE) :: MinDepmms, MinIcemms
   end type InputParameters

   type, bind(C) :: LocalParameters
      real(C_DOUBLE) :: tair_relax, VZ_relax, RH_relax
      integer(C_INT) :: couplingIndexI
      real(C_DOUBLE) :: couplingTsurf, lat, lon, sky_view
      integer(C_INT) :: InitLenI
   end type LocalParameters

   ! ---- Fortran views of the caller's arrays ---------------------------------------------------
   type :: InputArrays
      real(C_DOUBLE), pointer :: Tair(:) => null(), Tdew(:) => null(), VZ(:) => null(), Rhz(:) => null()
      real(C_DOUBLE), pointer :: prec(:) => null(), SW(:) => null(), LW(:) => null()
      real(C_DOUBLE), pointer :: SW_dir(:) => null(), LW_net(:) => null(), TSurfObs(:) => null()
      integer(C_INT), pointer :: PrecPhase(:) => null()
      real(C_DOUBLE), pointer :: local_horizons(:) => null(), depth(:) => null()
      integer(C_INT), pointer :: year(:) => null(), month(:) => null(), day(:) => null()
      integer(C_INT), pointer :: hour(:) => null(), minute(:) => null(), second(:) => null()
   end type InputArrays

   type :: OutputArrays
      real(C_DOUBLE), pointer :: TsurfOut(:) => null(), SnowOut(:) => null(), WaterOut(:) => null()
      real(C_DOUBLE), pointer :: IceOut(:) => null(), DepositOut(:) => null(), Ice2Out(:) => null()
   end type OutputArrays

   ! ---- state types: session handle + what a main reads between calls --------------------------
   !> roadsurf_session_open's handle.  Initialization stores the same handle in every state type it
   !! receives, because the step procedures each see a different subset of them.
   type :: ModelSettings
      type(C_PTR) :: session = C_NULL_PTR
      integer :: SimLen = 0, InitLenI = 0, NLayers = 0, coupling_minutes = 0, outputStep = 0
      logical :: use_coupling = .false., use_relaxation = .false., force_tsurf = .false.
      logical :: simulation_failed = .false.
      real(8) :: DTSecs = 0, Tph = 0, tsurfOutputDepth = 0, couplingEffectReduction = 0
   end type ModelSettings

   type :: SurfaceVariables
      type(C_PTR) :: session = C_NULL_PTR
      !> outputs of the most recent SaveOutput (TsurfAve and the five storages of that step)
      real(8) :: TsurfAve = 0, SrfWatmms = 0, SrfSnowmms = 0, SrfIcemms = 0, SrfIce2mms = 0, SrfDepmms = 0
   end type SurfaceVariables

   type :: GroundVariables
      type(C_PTR) :: session = C_NULL_PTR
      real(8) :: Albedo = 0       !< passed to CalcAlbedo by the main (a no-op here)
   end type GroundVariables

   type :: AtmVariables
      type(C_PTR) :: session = C_NULL_PTR
   end type AtmVariables

   type :: CouplingVariables
      type(C_PTR) :: session = C_NULL_PTR
      logical :: Coupling_failed = .false.   !< RS_ST_COUPLING_FAILED of the point, after CheckEndCoupling
   end type CouplingVariables

   type :: PhysicalParameters
      real(8) :: MaxPormms = 0    !< passed to RoadCond by the main
   end type PhysicalParameters

   type :: RoadCondParameters
      real(8) :: Snow2IceFac = 0  !< passed to WearFactors by the main
   end type RoadCondParameters

   type :: WearingFactors
      real(8) :: unused = 0
   end type WearingFactors

   type :: InputRadiationCoefficient   !< unused by the reference as well (src/InputRadiationCoefficient.f90.inc)
      real(8) :: unused = 0
   end type InputRadiationCoefficient

end module RoadSurfVariables
