!> ISO_C_BINDING interface to libroadsurf_b200.so for Fortran main programs.
!!
!! NOT COMPILED OR TESTED IN THIS REPOSITORY: no Fortran compiler exists in the build image
!! (see DESIGN.md section 1).  The interop types are the reference's own Bind(C) types
!! (src/InputPointers.f90.inc, OutputPointers.f90.inc, InputSettings.f90.inc,
!! InputParameters.f90.inc, LocalParameters.f90.inc), so a program that already fills them for
!! the reference's `runsimulation` (examples/example1/src/Simulation.f90:4-6) only changes what it
!! links against.
!!
!!   gfortran -c roadsurf_b200.f90 main.f90
!!   gfortran main.o roadsurf_b200.o -L<repo>/roadsurf_b200 -lroadsurf_b200
module RoadSurfB200
   use, intrinsic :: ISO_C_BINDING
   use RoadSurfVariables, only: InputPointers, OutputPointers, InputSettings, &
                                InputParameters, LocalParameters
   implicit none

   integer(C_INT), parameter :: RS_OK = 0
   integer(C_INT), parameter :: RS_ST_FAILED = 1, RS_ST_BAD_INPUT = 2, RS_ST_ABNORMAL_TSURF = 4, &
                                RS_ST_COUPLING_USED = 8, RS_ST_COUPLING_FAILED = 16, &
                                RS_ST_BL_NOT_CONVERGED = 32, RS_ST_SOLAR_GEOMETRY = 64

   interface
      !> The reference entry point, one point (include/roadsurf_b200.h: runsimulation).
      subroutine runsimulation(outPointers, inPointers, inSettings, inputParam, localParam) &
         bind(C, name="runsimulation")
         import :: OutputPointers, InputPointers, InputSettings, InputParameters, LocalParameters
         type(OutputPointers), intent(INOUT) :: outPointers
         type(InputPointers), intent(IN) :: inPointers
         type(InputSettings), intent(IN) :: inSettings
         type(InputParameters), intent(IN) :: inputParam
         type(LocalParameters), intent(IN) :: localParam
      end subroutine runsimulation

      !> Batched form: arrays of C pointers to the per-point structs (C_LOC of each element).
      integer(C_INT) function roadsurf_run_batch(npoints, outPtrs, inPtrs, settings, params, &
                                                 localPtrs, ngpus, status) &
         bind(C, name="roadsurf_run_batch")
         import :: C_INT, C_PTR, InputSettings, InputParameters
         integer(C_INT), value :: npoints
         type(C_PTR), intent(IN) :: outPtrs(*)     !< C_LOC(outPointers(p))
         type(C_PTR), intent(IN) :: inPtrs(*)      !< C_LOC(inPointers(p))
         type(InputSettings), intent(IN) :: settings
         type(InputParameters), intent(IN) :: params
         type(C_PTR), intent(IN) :: localPtrs(*)   !< C_LOC(localParam(p))
         integer(C_INT), value :: ngpus            !< <= 0: all visible devices
         integer(C_INT), intent(OUT) :: status(*)  !< RS_ST_* bit set per point
      end function roadsurf_run_batch

      integer(C_INT) function roadsurf_device_count() bind(C, name="roadsurf_device_count")
         import :: C_INT
      end function roadsurf_device_count

      type(C_PTR) function roadsurf_last_error() bind(C, name="roadsurf_last_error")
         import :: C_PTR
      end function roadsurf_last_error
   end interface

contains

   !> Convenience wrapper: run all points of Fortran arrays of the interop structs.
   subroutine RunSimulationBatch(outP, inP, settings, params, localP, status, rc)
      type(OutputPointers), intent(INOUT), target :: outP(:)
      type(InputPointers), intent(IN), target :: inP(:)
      type(InputSettings), intent(IN) :: settings
      type(InputParameters), intent(IN) :: params
      type(LocalParameters), intent(IN), target :: localP(:)
      integer(C_INT), intent(OUT) :: status(:)
      integer(C_INT), intent(OUT) :: rc
      type(C_PTR), allocatable :: po(:), pi(:), pl(:)
      integer :: p, n

      n = size(inP)
      allocate (po(n), pi(n), pl(n))
      do p = 1, n
         po(p) = C_LOC(outP(p))
         pi(p) = C_LOC(inP(p))
         pl(p) = C_LOC(localP(p))
      end do
      rc = roadsurf_run_batch(int(n, C_INT), po, pi, settings, params, pl, 0_C_INT, status)
   end subroutine RunSimulationBatch

end module RoadSurfB200
