!> @file
!! @brief Fortran interfaces (ISO_C_BINDING) of the library-management entry points of libipb200.so (include/ipb200.h,
!! "library management"): device selection, in-library multi-GPU sharding, page-locked host memory, plan cache.
!!
!! None of these exists in NCEPLIBS-ip; a caller that only wants the drop-in behaviour needs none of them.  They are what
!! makes the drop-in fast (INTEGRATION.md section 4):
!!
!!     use ipb200_gpu
!!     integer(c_int) :: n, rc
!!     n  = ipgpu_set_devices(ipgpu_device_count())             ! every later ipolates call shards km over all GPUs
!!     rc = ipgpu_host_register(c_loc(go), int(sizeof(go), c_long_long))   ! long-lived arrays: copied at link rate
!!
!! Source only, like ip_mod.F90 (no Fortran compiler in the build image); tests/test_abi_and_math.py checks the line lengths
!! and that every bind(c) name below is exported by the library.
module ipb200_gpu
  use iso_c_binding
  implicit none
  public
  interface
     !> CUDA devices visible to the process
     function ipgpu_device_count() bind(c, name='ipgpu_device_count') result(n)
       import :: c_int
       integer(c_int) :: n
     end function ipgpu_device_count
     !> cudaSetDevice for the calling thread; every ipolates / ipolatev call works on the device current when it enters
     function ipgpu_set_device(device) bind(c, name='ipgpu_set_device') result(rc)
       import :: c_int
       integer(c_int), value :: device
       integer(c_int) :: rc
     end function ipgpu_set_device
     !> host-array calls shard their field batch over the first n devices (n <= 1: the current one); returns the count in use
     function ipgpu_set_devices(n) bind(c, name='ipgpu_set_devices') result(used)
       import :: c_int
       integer(c_int), value :: n
       integer(c_int) :: used
     end function ipgpu_set_devices
     !> page-lock an array the caller owns; afterwards it is copied directly at link rate.  0, or the CUDA error code
     function ipgpu_host_register(p, bytes) bind(c, name='ipgpu_host_register') result(rc)
       import :: c_int, c_ptr, c_long_long
       type(c_ptr), value :: p
       integer(c_long_long), value :: bytes
       integer(c_int) :: rc
     end function ipgpu_host_register
     function ipgpu_host_unregister(p) bind(c, name='ipgpu_host_unregister') result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: p
       integer(c_int) :: rc
     end function ipgpu_host_unregister
     !> page-locked host memory on the NUMA node of the current device's PCIe link (use with c_f_pointer)
     function ipgpu_host_alloc(bytes) bind(c, name='ipgpu_host_alloc') result(p)
       import :: c_ptr, c_long_long
       integer(c_long_long), value :: bytes
       type(c_ptr) :: p
     end function ipgpu_host_alloc
     subroutine ipgpu_host_free(p) bind(c, name='ipgpu_host_free')
       import :: c_ptr
       type(c_ptr), value :: p
     end subroutine ipgpu_host_free
     !> drop every cached plan and the staging buffers, on every device
     subroutine ipgpu_plan_clear() bind(c, name='ipgpu_plan_clear')
     end subroutine ipgpu_plan_clear
     function ipgpu_plan_count() bind(c, name='ipgpu_plan_count') result(n)
       import :: c_long_long
       integer(c_long_long) :: n
     end function ipgpu_plan_count
     function ipgpu_plan_bytes() bind(c, name='ipgpu_plan_bytes') result(n)
       import :: c_long_long
       integer(c_long_long) :: n
     end function ipgpu_plan_bytes
     !> LRU capacity of the per-device plan cache (default 24 GiB)
     subroutine ipgpu_set_plan_cache_bytes(bytes) bind(c, name='ipgpu_set_plan_cache_bytes')
       import :: c_long_long
       integer(c_long_long), value :: bytes
     end subroutine ipgpu_set_plan_cache_bytes
     !> staging size per in-flight chunk of the host-array calls (default 768 MiB)
     subroutine ipgpu_set_chunk_bytes(bytes) bind(c, name='ipgpu_set_chunk_bytes')
       import :: c_long_long
       integer(c_long_long), value :: bytes
     end subroutine ipgpu_set_chunk_bytes
     !> CUDA-event milliseconds of the last call: plan build (0 on a cache hit), apply stage, and whether the plan was cached
     subroutine ipgpu_last_timing(plan_ms, apply_ms, plan_cache_hit) bind(c, name='ipgpu_last_timing')
       import :: c_double, c_int
       real(c_double), intent(out) :: plan_ms, apply_ms
       integer(c_int), intent(out) :: plan_cache_hit
     end subroutine ipgpu_last_timing
     !> measured ceiling of the host link of the current device, GB/s, with copies of the given size
     subroutine ipgpu_link_probe(bytes, h2d_gbs, d2h_gbs, both_gbs) bind(c, name='ipgpu_link_probe')
       import :: c_double, c_long_long
       integer(c_long_long), value :: bytes
       real(c_double), intent(out) :: h2d_gbs, d2h_gbs, both_gbs
     end subroutine ipgpu_link_probe
  end interface
end module ipb200_gpu
