!> Module RoadSurf for the B200 library: the 14 public procedures of the reference's module RoadSurf
!! (src/RoadSurf.f90:257-270) with their argument lists (src/RoadSurf.f90:9-252), forwarding through
!! ISO_C_BINDING to the step-granular C entry points of libroadsurf_b200.so
!! (include/roadsurf_b200.h: roadsurf_session_open / roadsurf_step / roadsurf_session_fetch).
!!
!! NOT COMPILED OR TESTED IN THIS REPOSITORY (no Fortran compiler in the build image); the C entry
!! points it binds are tested (tests/test_gpu_parity.py::test_stepwise_*).  Build sketch:
!!   gfortran -c RoadSurfVariables.f90 RoadSurf.f90 Simulation.f90
!!   g++ main.o Simulation.o RoadSurf.o RoadSurfVariables.o -L<repo>/roadsurf_b200 -lroadsurf_b200 -lgfortran
!!
!! How the per-step call sequence of a main (examples/example1/src/Simulation.f90:58-115) maps:
!!
!!   ConnectFortran2Carrays    host only: C_F_POINTER association, as in the reference
!!   Initialization            roadsurf_session_open: inputs uploaded once, state planes allocated,
!!                             outputs pre-filled with -9999.0; settings / handles filled in
!!   CheckValues               no-op (fused).  The in-kernel range checks set the point's status; the
!!   CouplingOperations1       no-op (fused; the kernel rewinds inside the launch, `i` is never changed)
!!   SetCurrentValues          no-op (fused)
!!   RelaxationOperations      no-op (fused)
!!   PrecipitationToStorage    no-op (fused)
!!   ModRadiationBySurroundings no-op (fused; the caller's SW / SW_dir / LW arrays are NOT rewritten)
!!   BalanceModelOneStep       no-op (fused)
!!   WearFactors, RoadCond, CalcAlbedo   no-ops (fused)
!!   SaveOutput(i)             roadsurf_step(session, i): ONE launch of the step kernel that performs all
!!                             of the above for step i (step_begin = step_end = i, state resident on the
!!                             device), then roadsurf_session_fetch: out(i) into the caller's arrays
!!   CheckEndCoupling          reads the point's status word: settings%simulation_failed is set here, so
!!                             the main's `do while (... .not. simulation_failed)` ends as in the reference
!!   lastValues (external)     no-op: the kernel runs the last-value step when SaveOutput(SimLen) asks for it
!!
!! Consequences a caller should know: (1) state components other than the ones RoadSurfVariables keeps
!! are not visible between calls; (2) a coupling window is executed as a whole by the SaveOutput call
!! that enters it (the window's outputs appear at that moment, later SaveOutput calls inside it launch
!! nothing); (3) RS_SESSION_CHUNK > 1 lets a launch run ahead (results unchanged).
module RoadSurf
   use, intrinsic :: ISO_C_BINDING
   use RoadSurfVariables
   implicit none
   private

   public :: ConnectFortran2Carrays, Initialization, CheckValues, CouplingOperations1, RelaxationOperations
   public :: SetCurrentValues, BalanceModelOneStep, SaveOutput, CheckEndCoupling, PrecipitationToStorage
   public :: ModRadiationBySurroundings, WearFactors, RoadCond, CalcAlbedo

   !> steps a launch may run ahead of the step that was asked for (1 = strictly one step per launch)
   integer(C_INT), public :: RS_SESSION_CHUNK = 120

   integer(C_INT), parameter :: RS_ST_FAILED = 1, RS_ST_COUPLING_FAILED = 16

   interface
      integer(C_INT) function roadsurf_session_open(npoints, outPtrs, inPtrs, settings, params, localPtrs, session) &
         bind(C, name="roadsurf_session_open")
         import :: C_INT, C_PTR, InputSettings, InputParameters
         integer(C_INT), value :: npoints
         type(C_PTR), intent(IN) :: outPtrs(*), inPtrs(*), localPtrs(*)
         type(InputSettings), intent(IN) :: settings
         type(InputParameters), intent(IN) :: params
         type(C_PTR), intent(OUT) :: session
      end function roadsurf_session_open
      integer(C_INT) function roadsurf_step(session, i) bind(C, name="roadsurf_step")
         import :: C_INT, C_PTR
         type(C_PTR), value :: session
         integer(C_INT), value :: i
      end function roadsurf_step
      integer(C_INT) function roadsurf_session_fetch(session, i, status) bind(C, name="roadsurf_session_fetch")
         import :: C_INT, C_PTR
         type(C_PTR), value :: session
         integer(C_INT), value :: i
         integer(C_INT), intent(OUT) :: status(*)
      end function roadsurf_session_fetch
      integer(C_INT) function roadsurf_session_set_chunk(session, steps) bind(C, name="roadsurf_session_set_chunk")
         import :: C_INT, C_PTR
         type(C_PTR), value :: session
         integer(C_INT), value :: steps
      end function roadsurf_session_set_chunk
      subroutine roadsurf_session_close(session) bind(C, name="roadsurf_session_close")
         import :: C_PTR
         type(C_PTR), value :: session
      end subroutine roadsurf_session_close
   end interface

contains

   subroutine ConnectFortran2Carrays(inPointers, modelInput, outPointers, modelOutput)
      type(InputPointers), intent(IN) :: inPointers
      type(OutputPointers), intent(INOUT) :: outPointers
      type(InputArrays), intent(OUT) :: modelInput
      type(OutputArrays), intent(OUT) :: modelOutput
      integer :: n
      n = inPointers%inputLen
      call C_F_POINTER(inPointers%c_tair, modelInput%Tair, [n])
      call C_F_POINTER(inPointers%c_tdew, modelInput%Tdew, [n])
      call C_F_POINTER(inPointers%c_VZ, modelInput%VZ, [n])
      call C_F_POINTER(inPointers%c_Rhz, modelInput%Rhz, [n])
      call C_F_POINTER(inPointers%c_prec, modelInput%prec, [n])
      call C_F_POINTER(inPointers%c_SW, modelInput%SW, [n])
      call C_F_POINTER(inPointers%c_LW, modelInput%LW, [n])
      call C_F_POINTER(inPointers%c_SW_dir, modelInput%SW_dir, [n])
      call C_F_POINTER(inPointers%c_LW_net, modelInput%LW_net, [n])
      call C_F_POINTER(inPointers%c_TSurfObs, modelInput%TSurfObs, [n])
      call C_F_POINTER(inPointers%c_PrecPhase, modelInput%PrecPhase, [n])
      call C_F_POINTER(inPointers%c_local_horizons, modelInput%local_horizons, [360])
      call C_F_POINTER(inPointers%c_Depth, modelInput%depth, [n])
      call C_F_POINTER(inPointers%c_year, modelInput%year, [n])
      call C_F_POINTER(inPointers%c_month, modelInput%month, [n])
      call C_F_POINTER(inPointers%c_day, modelInput%day, [n])
      call C_F_POINTER(inPointers%c_hour, modelInput%hour, [n])
      call C_F_POINTER(inPointers%c_minute, modelInput%minute, [n])
      call C_F_POINTER(inPointers%c_second, modelInput%second, [n])
      n = outPointers%outputLen
      call C_F_POINTER(outPointers%c_TsurfOut, modelOutput%TsurfOut, [n])
      call C_F_POINTER(outPointers%c_SnowOut, modelOutput%SnowOut, [n])
      call C_F_POINTER(outPointers%c_WaterOut, modelOutput%WaterOut, [n])
      call C_F_POINTER(outPointers%c_IceOut, modelOutput%IceOut, [n])
      call C_F_POINTER(outPointers%c_DepositOut, modelOutput%DepositOut, [n])
      call C_F_POINTER(outPointers%c_Ice2Out, modelOutput%Ice2Out, [n])
   end subroutine ConnectFortran2Carrays

   !> Opens the device session for this point.  The C structs the session needs are rebuilt from the
   !! associated arrays (Initialization does not receive the InputPointers / OutputPointers).
   subroutine Initialization(modelInput, inSettings, settings, modelOutput, atm, surf, inputParam, &
                             localParam, coupling, phy, ground, condParam)
      type(InputSettings), intent(IN) :: inSettings
      type(InputParameters), intent(IN) :: inputParam
      type(LocalParameters), intent(IN), target :: localParam
      type(InputArrays), intent(INOUT) :: modelInput    ! (the reference says OUT and relies on it staying associated)
      type(OutputArrays), intent(INOUT) :: modelOutput
      type(AtmVariables), intent(OUT) :: atm
      type(CouplingVariables), intent(OUT) :: coupling
      type(ModelSettings), intent(OUT) :: settings
      type(PhysicalParameters), intent(OUT) :: phy
      type(GroundVariables), intent(OUT) :: ground
      type(SurfaceVariables), intent(OUT) :: surf
      type(RoadCondParameters), intent(OUT) :: condParam
      type(InputPointers), target :: ip
      type(OutputPointers), target :: op
      type(C_PTR) :: pin(1), pout(1), ploc(1), session
      integer(C_INT) :: rc

      ip%inputLen = size(modelInput%Tair)
      ip%c_tair = C_LOC(modelInput%Tair(1));     ip%c_tdew = C_LOC(modelInput%Tdew(1))
      ip%c_VZ = C_LOC(modelInput%VZ(1));         ip%c_Rhz = C_LOC(modelInput%Rhz(1))
      ip%c_prec = C_LOC(modelInput%prec(1));     ip%c_SW = C_LOC(modelInput%SW(1))
      ip%c_LW = C_LOC(modelInput%LW(1));         ip%c_SW_dir = C_LOC(modelInput%SW_dir(1))
      ip%c_LW_net = C_LOC(modelInput%LW_net(1)); ip%c_TSurfObs = C_LOC(modelInput%TSurfObs(1))
      ip%c_PrecPhase = C_LOC(modelInput%PrecPhase(1))
      ip%c_local_horizons = C_LOC(modelInput%local_horizons(1))
      ip%c_Depth = C_LOC(modelInput%depth(1))
      ip%c_year = C_LOC(modelInput%year(1));     ip%c_month = C_LOC(modelInput%month(1))
      ip%c_day = C_LOC(modelInput%day(1));       ip%c_hour = C_LOC(modelInput%hour(1))
      ip%c_minute = C_LOC(modelInput%minute(1)); ip%c_second = C_LOC(modelInput%second(1))
      op%outputLen = size(modelOutput%TsurfOut)
      op%c_TsurfOut = C_LOC(modelOutput%TsurfOut(1));     op%c_SnowOut = C_LOC(modelOutput%SnowOut(1))
      op%c_WaterOut = C_LOC(modelOutput%WaterOut(1));     op%c_IceOut = C_LOC(modelOutput%IceOut(1))
      op%c_DepositOut = C_LOC(modelOutput%DepositOut(1)); op%c_Ice2Out = C_LOC(modelOutput%Ice2Out(1))
      pin(1) = C_LOC(ip); pout(1) = C_LOC(op); ploc(1) = C_LOC(localParam)

      settings%SimLen = inSettings%SimLen
      settings%InitLenI = localParam%InitLenI
      settings%NLayers = inSettings%NLayers
      settings%DTSecs = inSettings%DTSecs
      settings%Tph = inSettings%DTSecs/3600.0_8
      settings%tsurfOutputDepth = inSettings%tsurfOutputDepth
      settings%coupling_minutes = inSettings%coupling_minutes
      settings%couplingEffectReduction = inSettings%couplingEffectReduction
      settings%outputStep = inSettings%outputStep
      ! the library decides per point whether coupling / relaxation really run (missing observation
      ! or targets switch them off, src/InputOutput.f90:23-36); the flags only steer the main's calls
      settings%use_coupling = inSettings%use_coupling == 1
      settings%use_relaxation = inSettings%use_relaxation == 1
      settings%force_tsurf = inSettings%force_tsurf == 1
      phy%MaxPormms = inputParam%MaxPormms
      ground%Albedo = inputParam%Albedo
      condParam%Snow2IceFac = inputParam%Snow2IceFac

      rc = roadsurf_session_open(1_C_INT, pout, pin, inSettings, inputParam, ploc, session)
      settings%simulation_failed = rc /= 0
      if (rc == 0) rc = roadsurf_session_set_chunk(session, RS_SESSION_CHUNK)
      settings%session = session
      surf%session = session
      ground%session = session
      atm%session = session
      coupling%session = session
   end subroutine Initialization

   subroutine CheckValues(modelInput, i, settings, surf, localParam)
      type(InputArrays), intent(INOUT) :: modelInput
      integer, intent(IN) :: i
      type(SurfaceVariables), intent(IN) :: surf
      type(ModelSettings), intent(INOUT) :: settings
      type(LocalParameters), intent(IN) :: localParam
   end subroutine CheckValues

   subroutine CouplingOperations1(i, coupling, surf, settings, ground, modelInput, CP, localParam)
      type(ModelSettings), intent(IN) :: settings
      type(InputArrays), intent(INOUT) :: modelInput
      type(RoadCondParameters), intent(IN) :: CP
      integer, intent(INOUT) :: i
      type(CouplingVariables), intent(INOUT) :: coupling
      type(SurfaceVariables), intent(INOUT) :: surf
      type(GroundVariables), intent(INOUT) :: ground
      type(LocalParameters), intent(IN) :: localParam
   end subroutine CouplingOperations1

   subroutine RelaxationOperations(i, atm, settings, ground)
      integer, intent(IN) :: i
      type(ModelSettings), intent(IN) :: settings
      type(AtmVariables), intent(INOUT) :: atm
      type(GroundVariables), intent(INOUT) :: ground
   end subroutine RelaxationOperations

   subroutine SetCurrentValues(i, modelInput, atm, settings, surf, coupling, ground)
      integer, intent(IN) :: i
      type(ModelSettings), intent(IN) :: settings
      type(InputArrays), intent(IN) :: modelInput
      type(CouplingVariables), intent(IN) :: coupling
      type(AtmVariables), intent(INOUT) :: atm
      type(SurfaceVariables), intent(INOUT) :: surf
      type(GroundVariables), intent(INOUT) :: ground
   end subroutine SetCurrentValues

   subroutine BalanceModelOneStep(SWi, LWi, phy, ground, surf, atm, settings, coupling, modelInput, inputIdx, condParam)
      real(8), intent(IN) :: SWi, LWi
      type(PhysicalParameters), intent(INOUT) :: phy
      type(CouplingVariables), intent(IN) :: coupling
      type(InputArrays), intent(IN) :: modelInput
      type(GroundVariables), intent(INOUT) :: ground
      type(SurfaceVariables), intent(INOUT) :: surf
      type(AtmVariables), intent(INOUT) :: atm
      type(ModelSettings), intent(INOUT) :: settings
      type(RoadCondParameters), intent(IN) :: condParam
      integer, intent(IN) :: inputIdx
   end subroutine BalanceModelOneStep

   !> The one procedure that launches: everything the reference does for step i, in one kernel launch.
   subroutine SaveOutput(modelOutput, i, surf)
      integer, intent(IN) :: i
      type(SurfaceVariables), intent(IN) :: surf
      type(OutputArrays), intent(INOUT) :: modelOutput
      integer(C_INT) :: rc, status(1)
      if (.not. C_ASSOCIATED(surf%session)) return
      rc = roadsurf_step(surf%session, int(i, C_INT))
      if (rc == 0) rc = roadsurf_session_fetch(surf%session, int(i, C_INT), status)
   end subroutine SaveOutput

   subroutine CheckEndCoupling(i, settings, coupling, surf)
      integer, intent(IN) :: i
      type(ModelSettings), intent(IN) :: settings   ! (IN in the reference; the failure flag is set through a pointer)
      type(SurfaceVariables), intent(INOUT) :: surf
      type(CouplingVariables), intent(INOUT) :: coupling
      integer(C_INT) :: rc, status(1)
      if (.not. C_ASSOCIATED(surf%session)) return
      rc = roadsurf_session_fetch(surf%session, int(i, C_INT), status)
      coupling%Coupling_failed = iand(status(1), RS_ST_COUPLING_FAILED) /= 0
      call set_failed(settings, rc /= 0 .or. iand(status(1), RS_ST_FAILED) /= 0)
   contains
      !> settings is intent(IN) in the reference's interface, which a main's loop condition nevertheless
      !! relies on seeing updated (the reference sets it in CheckValues, where it is INOUT).
      subroutine set_failed(s, failed)
         type(ModelSettings), intent(IN), target :: s
         logical, intent(IN) :: failed
         type(ModelSettings), pointer :: p
         if (.not. failed) return
         call C_F_POINTER(C_LOC(s), p)
         p%simulation_failed = .true.
      end subroutine set_failed
   end subroutine CheckEndCoupling

   subroutine PrecipitationToStorage(settings, CP, PrecPhase, atm, surf)
      type(ModelSettings), intent(IN) :: settings
      type(RoadCondParameters), intent(IN) :: CP
      integer, intent(IN) :: PrecPhase
      type(AtmVariables), intent(INOUT) :: atm
      type(SurfaceVariables), intent(INOUT) :: surf
   end subroutine PrecipitationToStorage

   subroutine ModRadiationBySurroundings(modelInput, inputParam, localParam, i)
      type(InputArrays), intent(INOUT) :: modelInput
      type(InputParameters), intent(IN) :: inputParam
      type(LocalParameters), intent(IN) :: localParam
      integer, intent(IN) :: i
   end subroutine ModRadiationBySurroundings

   subroutine WearFactors(Snow2IceFac, Tph, surf, wearF)
      real(8), intent(IN) :: Tph
      type(SurfaceVariables), intent(IN) :: surf
      type(WearingFactors), intent(OUT) :: wearF
      real(8), intent(INOUT) :: Snow2IceFac
      wearF%unused = 0
   end subroutine WearFactors

   subroutine RoadCond(MaxPormms, surf, atm, settings, CP, wearF)
      real(8), intent(IN) :: MaxPormms
      type(ModelSettings), intent(IN) :: settings
      type(RoadCondParameters), intent(INOUT) :: CP
      type(SurfaceVariables), intent(INOUT) :: surf
      type(AtmVariables), intent(INOUT) :: atm
      type(WearingFactors), intent(IN) :: wearF
   end subroutine RoadCond

   subroutine CalcAlbedo(albedo, surf, cp)
      type(SurfaceVariables), intent(IN) :: surf
      type(RoadCondParameters), intent(IN) :: cp
      real(8), intent(INOUT) :: albedo
   end subroutine CalcAlbedo

end module RoadSurf

!> External in the reference as well (src/InputOutput.f90:169-198; called by Simulation.f90:105).  The
!! kernel performs the last-value step itself when SaveOutput(SimLen) asks for it.
subroutine lastValues(modelInput, atm, settings, ground, surf)
   use RoadSurfVariables
   implicit none
   type(ModelSettings), intent(IN) :: settings
   type(InputArrays), intent(IN) :: modelInput
   type(AtmVariables), intent(INOUT) :: atm
   type(GroundVariables), intent(INOUT) :: ground
   type(SurfaceVariables), intent(INOUT) :: surf
end subroutine lastValues
