!> ISO_C_BINDING interfaces of libnm_b200.so for the Fortran host of NormalModes.
!!
!! The 18 pEVSL entry points the reference already calls (pevsl_*_f90, src/mod_matvec.f90 and
!! src/mod_pevsl.f90) need NO interface: they are external procedures with the same names and
!! by-reference arguments, resolved by linking against -lnm_b200 instead of -lpevsl.
!! This module declares what is NEW (include/nm_b200.h, include/pevsl_f90.h):
!!   * runtime / communicator hand-over         (INTEGRATION.md section 1)
!!   * device-resident operator registration    (INTEGRATION.md section 3, "level 1")
!!   * device FE assembly                       (INTEGRATION.md section 4)
!! No Fortran compiler exists in the build image: this file is shipped as text and has not been
!! compiled here; the same call sequences are exercised from C types in
!! tests/test_gpu_parity.py::test_f90_abi_device_resident_operators_fluid_solid.
module nm_b200_iface
  use iso_c_binding
  implicit none

  interface
     ! ---- runtime: one process (MPI rank) per GPU ----------------------------------------------
     integer(c_int) function nm_init(device) bind(C, name="nm_init")
       import :: c_int
       integer(c_int), value :: device            ! < 0: keep the current CUDA device
     end function nm_init

     integer(c_int) function nm_comm_unique_id(id128) bind(C, name="nm_comm_unique_id")
       import :: c_int, c_char
       character(kind=c_char) :: id128(128)       ! rank 0 creates it, MPI_Bcast distributes it
     end function nm_comm_unique_id

     integer(c_int) function nm_comm_init(rank, nranks, id128) bind(C, name="nm_comm_init")
       import :: c_int, c_char
       integer(c_int), value :: rank, nranks
       character(kind=c_char) :: id128(128)
     end function nm_comm_init

     integer(c_int) function nm_comm_finalize() bind(C, name="nm_comm_finalize")
       import :: c_int
     end function nm_comm_finalize

     integer(c_int) function nm_last_error() bind(C, name="nm_last_error")
       import :: c_int
     end function nm_last_error

     ! ---- Jacobi scaling on the device (replaces Bdiagscaling / Apdiagscaling, optional) ---------
     integer(c_int) function nm_parcsr_jacobi_scale(mat, sgn, d_host) bind(C, name="nm_parcsr_jacobi_scale")
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: mat                  ! handle returned by PEVSL_PARCSRCREATE_F90
       real(c_double), value :: sgn               ! +1 for B, -1 for Ap (Ap := -CGM%Ap)
       real(c_double) :: d_host(*)                ! receives d = 1/sqrt(diag), local rows
     end function nm_parcsr_jacobi_scale

     ! ---- FE assembly: cg_create_matrix (pattern on the host, values on the device) -------------
     integer(c_int) function nm_fem_create(ntet, nvert, ele, neigh, node, porder, vs, nproc, part, rank, fem) &
          bind(C, name="nm_fem_create")
       import :: c_int, c_double, c_ptr
       integer(c_int), value :: ntet, nvert, porder, nproc, rank
       integer(c_int) :: ele(4, *), neigh(4, *)   ! 0-based vertex / element ids, -1 = boundary face
       real(c_double) :: node(3, *), vs(*)        ! vs(pNp, ntet): fluid element <=> max(vs) < 1e-6
       integer(c_int) :: part(*)                  ! owner rank of every node (ParMETIS result), size nn
       type(c_ptr) :: fem                         ! out
     end function nm_fem_create

     integer(c_int) function nm_fem_assemble_values(fem, job, vp, vs, rho, g0) bind(C, name="nm_fem_assemble_values")
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: fem
       integer(c_int), value :: job
       real(c_double) :: vp(*), vs(*), rho(*), g0(*)   ! (pNp, ntet); g0 (3, pNp, ntet) in m/s^2, JOB >= 2
     end function nm_fem_assemble_values

     integer(c_int) function nm_fem_matrix_sizes(fem, which, present, nrow_local, nnz_local) &
          bind(C, name="nm_fem_matrix_sizes")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: fem
       integer(c_int), value :: which             ! 0 A|Ad, 1 B, 2 E, 3 ET, 4 Ap
       integer(c_int) :: present, nrow_local
       integer(c_long_long) :: nnz_local
     end function nm_fem_matrix_sizes

     integer(c_int) function nm_fem_matrix_get(fem, which, rowdist, coldist, ia, ja, val) bind(C, name="nm_fem_matrix_get")
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: fem
       integer(c_int), value :: which
       integer(c_int) :: rowdist(*), coldist(*), ia(*), ja(*)   ! ja: 0-based GLOBAL column ids
       real(c_double) :: val(*)
     end function nm_fem_matrix_get

     integer(c_int) function nm_fem_free(fem) bind(C, name="nm_fem_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: fem
     end function nm_fem_free
  end interface

  ! ---- device-resident operator registration (by-reference, Fortran external convention) -------
  ! These follow the pEVSL naming, so they are called like the pevsl_*_f90 routines:
  !   call NM_SETBMV_PARCSR_F90(pevslAB, mymatvec%sBV)
  !   call NM_SETAMV_SOLID_F90(pevslAB, mymatvec%sAV, mymatvec%B%diag)
  !   call NM_SETAMV_FLUIDSOLID_F90(pevslAB, mymatvec%sAdV, mymatvec%sEV, mymatvec%sETV, &
  !                                 mymatvec%chebAp, mymatvec%B%diag, mymatvec%Ap%diag)
  !   call NM_SETAMV_PARCSR_F90(pevslB, mymatvec%sBV)        ! bounds of B~ / Ap~ in setupmatvec
  ! (handles are the integer*8 values PEVSL_PARCSRCREATE_F90 / pEVSL_SETUP_CHEBITER_F90 returned).
  external :: nm_setbmv_parcsr_f90, nm_setamv_parcsr_f90, nm_setamv_solid_f90, nm_setamv_fluidsolid_f90

contains

  !> mainnm.f90, right after mpi_init: bind this rank to a GPU of the box and create the library's
  !! NCCL communicator + NVLink peer window.  `bcast128` must broadcast 128 characters from rank 0
  !! (e.g. a wrapper of mpi_bcast(id, 128, mpi_character, 0, comm, ierr)).
  subroutine nm_b200_start(myrank, nproc, gpus_per_node, bcast128, ierr)
    integer, intent(in) :: myrank, nproc, gpus_per_node
    interface
       subroutine bcast128(buf)
         import :: c_char
         character(kind=c_char) :: buf(128)
       end subroutine bcast128
    end interface
    integer, intent(out) :: ierr
    character(kind=c_char) :: id(128)
    id = c_null_char
    ierr = nm_init(int(mod(myrank, gpus_per_node), c_int))
    if (ierr /= 0) return
    if (myrank == 0) ierr = nm_comm_unique_id(id)
    call bcast128(id)
    ierr = nm_comm_init(int(myrank, c_int), int(nproc, c_int), id)
  end subroutine nm_b200_start

end module nm_b200_iface
