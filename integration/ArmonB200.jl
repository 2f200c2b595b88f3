# ArmonB200.jl -- the package extension a maintainer of Armon.jl adds as `ext/ArmonB200.jl` to select the
# B200-native backend with `ArmonParameters(; use_gpu=true, device=:B200, ...)`.
#
# NOT EXECUTED IN THIS REPOSITORY: neither the build container nor the GPU box has a `julia` binary.  The same
# call sequence, through the same exported symbols, is exercised by the Python/ctypes twin
# (`armon.jl_b200/backend.py`, `blocks.py`, `solver.py`) and by `tests/test_gpu_parity.py`.
#
# It follows the two extension seams of the reference (SURVEY.md section 8b):
#   1. backend registration hooks   -- modelled on ext/ArmonCUDA.jl:9-24 and ext/ArmonKokkos.jl:83-151,183-185
#   2. kernel / step overloads on `ArmonParameters{T, <:B200Device}` -- the way ext/ArmonKokkos.jl:212-258
#      overrides `dtCFL_kernel` / `conservation_vars`
# Every `ccall` targets a declaration of include/armon_b200.h.
module ArmonB200

using Armon
import Armon: ArmonParameters, BlockGrid, LocalTaskBlock, SolverState, SolverException, solver_error,
              Axis, Side, split_axes

const LIB = get(ENV, "ARMON_B200_LIB", "libarmon_b200")

# ------------------------------------------------------------------------------------------------------------
# Errors: int status + armon_last_error()  ->  SolverException (src/utils.jl:102-117)
# ------------------------------------------------------------------------------------------------------------
const ERR_CATEGORY = Dict(1 => :config, 2 => :cpp, 3 => :cpp, 4 => :time, 5 => :config, 6 => :cpp)

function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:armon_last_error, LIB), Cstring, ()))
    solver_error(get(ERR_CATEGORY, Int(status), :cpp), msg)
end

macro b200call(sym, argtypes, args...)
    esc(:(check(ccall(($sym, LIB), Cint, $argtypes, $(args...)))))
end

# ------------------------------------------------------------------------------------------------------------
# Seam 1: device, arrays, fence
# ------------------------------------------------------------------------------------------------------------
mutable struct B200Device
    ctx::Ptr{Cvoid}
    function B200Device(id::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        @b200call(:armon_ctx_create, (Cint, Ptr{Ptr{Cvoid}}), id, ref)
        dev = new(ref[])
        finalizer(d -> ccall((:armon_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), d.ctx), dev)
    end
end

# Minimal device vector: what BlockData needs is `A{T,1}(undef, n)`, `length`, `copyto!` both ways
# (src/blocking/blocks.jl:36-43,121-143).  The Julia object owns the allocation; the library never frees it.
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    ptr::Ptr{T}
    len::Int
    dev::B200Device
end

const CURRENT_DEVICE = Ref{B200Device}()

function B200Array{Float64, 1}(::UndefInitializer, n::Integer)
    dev = CURRENT_DEVICE[]
    ref = Ref{Ptr{Float64}}(C_NULL)
    @b200call(:armon_alloc, (Ptr{Cvoid}, UInt64, Ptr{Ptr{Float64}}), dev.ctx, n, ref)
    arr = B200Array{Float64, 1}(ref[], n, dev)
    finalizer(a -> ccall((:armon_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), a.dev.ctx, a.ptr), arr)
end
Base.size(a::B200Array) = (a.len,)
Base.length(a::B200Array) = a.len
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where {T} = a.ptr
Base.getindex(::B200Array, _...) = error("scalar indexing of a B200Array: copy it to the host first")

function Base.copyto!(dst::B200Array{Float64}, src::Array{Float64})
    @b200call(:armon_copy_h2d, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, UInt64), dst.dev.ctx, dst, src, length(src))
    return dst
end
function Base.copyto!(dst::Array{Float64}, src::B200Array{Float64})
    @b200call(:armon_copy_d2h, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, UInt64), src.dev.ctx, dst, src, length(src))
    return dst
end

function Armon.create_device(::Val{:B200})
    flt = ccall((:armon_flt_size, LIB), Cint, ())       # ABI self-check, cf. ext/ArmonKokkos.jl:122-140
    idx = ccall((:armon_idx_size, LIB), Cint, ())
    abi = ccall((:armon_b200_abi_version, LIB), Cint, ())
    (flt == 8 && idx == 8 && abi == 2) ||
        solver_error(:config, "libarmon_b200 ABI mismatch (flt=$flt, idx=$idx, abi=$abi): this file binds ABI version 2")
    CURRENT_DEVICE[] = B200Device()
end
Armon.device_array_type(::B200Device) = B200Array
Armon.host_array_type(::B200Device) = Array

# backend_options: the fused solver handle + what it was built from
mutable struct B200Options
    solver::Ptr{Cvoid}
    math_mode::Cint      # ARMON_MATH_*: 0 strict (bit-exact vs the CPU path compiled without @fastmath), 1 fast, 2 ieee
    kernel_variant::Cint # ARMON_KERNEL_*: 0 auto
    cuda_graph::Cint     # 0 auto (grids <= 512^2 on one rank), 1 on, 2 off
    dirty::Bool          # device state not yet brought back to the canonical layout (armon_solver_finalize)
end

const KERNEL_VARIANTS = Dict(:auto => 0, :single => 1, :async => 4, :async2 => 5, :tma => 6)
const CUDA_GRAPH_MODES = Dict(:auto => 0, :on => 1, :off => 2)

function Armon.init_backend(params::ArmonParameters, ::B200Device;
                            math_mode = :strict, kernel_variant = :auto, cuda_graph = :auto, options...)
    params.backend_options = B200Options(C_NULL, math_mode === :fast ? 1 : math_mode === :ieee ? 2 : 0,
                                         KERNEL_VARIANTS[kernel_variant], CUDA_GRAPH_MODES[cuda_graph], false)
    return options
end

# Every reader of the bound arrays (device_to_host!, conservation_vars, dtCFL_kernel, the per-step kernels) must first
# bring the fused solver's rotating, possibly transposed buffers back to the canonical BlockData layout (+ stale p, c, g).
function ensure_canonical(params::ArmonParameters)
    opts = params.backend_options
    if opts.dirty && opts.solver != C_NULL
        @b200call(:armon_solver_finalize, (Ptr{Cvoid},), opts.solver)
        opts.dirty = false
    end
    return nothing
end

Base.wait(params::ArmonParameters{<:Any, <:B200Device}) =
    @b200call(:armon_ctx_sync, (Ptr{Cvoid},), params.device.ctx)

function Armon.device_memory_info(dev::B200Device)
    free, total = Ref{UInt64}(0), Ref{UInt64}(0)
    @b200call(:armon_device_memory_info, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}), dev.ctx, free, total)
    return (total = total[], free = free[])
end

function Armon.print_device_info(io::IO, pad::Int, p::ArmonParameters{<:Any, <:B200Device})
    buf = Vector{UInt8}(undef, 256)
    @b200call(:armon_device_name, (Ptr{Cvoid}, Ptr{UInt8}, Cint), p.device.ctx, buf, length(buf))
    Armon.print_parameter(io, pad, "GPU", true, nl = false)
    println(io, ": ", unsafe_string(pointer(buf)), " (libarmon_b200, fused axis sweeps)")
end

# ------------------------------------------------------------------------------------------------------------
# Mirror structs of include/armon_b200.h (isbits, C layout)
# ------------------------------------------------------------------------------------------------------------
struct CDims;   nx::Int64; ny::Int64; g::Int64; end
struct CDomain; ix0::Int64; ix1::Int64; iy0::Int64; iy1::Int64; end
struct CTestCase
    test::Int32; eos::Int32
    high_rho::Float64; low_rho::Float64; high_E::Float64; low_E::Float64
    high_u::Float64; low_u::Float64; high_v::Float64; low_v::Float64
    sedov_r::Float64; gamma::Float64
    bc_u::NTuple{4, Float64}; bc_v::NTuple{4, Float64}
end
struct CSolverDesc
    dims::CDims
    global_nx::Int64; global_ny::Int64; origin_ix::Int64; origin_iy::Int64
    domain_size::NTuple{2, Float64}; origin::NTuple{2, Float64}
    riemann::Int32; limiter::Int32; projection::Int32; splitting::Int32
    cfl::Float64; maxtime::Float64; maxcycle::Int64
    cst_dt::Int32; Dt::Float64
    neighbours::NTuple{4, Int32}
    math_mode::Int32; march_segment::Int32; kernel_variant::Int32; cuda_graph::Int32
    tc::CTestCase
end
struct CTimeState
    cycle::Int64; time::Float64; current_dt::Float64; next_cycle_dt::Float64; error::Int32; done::Int32
    error_cycle::Int64
end
struct CCycleDiag   # armon_cycle_diag: one line of the `silent <= 1` log (src/solver.jl:359-371), produced on the device
    cycle::Int64; time::Float64; dt::Float64; mass::Float64; energy::Float64
end

test_code(::Armon.Sod) = 0; test_code(::Armon.Sod_y) = 1; test_code(::Armon.Sod_circ) = 2
test_code(::Armon.Bizarrium) = 3; test_code(::Armon.Sedov) = 4; test_code(::Armon.DebugIndexes) = 5
limiter_code(::Armon.NoLimiter) = 0; limiter_code(::Armon.MinmodLimiter) = 1; limiter_code(::Armon.SuperbeeLimiter) = 2
riemann_code(::Armon.RiemannGodunov) = 0; riemann_code(::Armon.RiemannGAD) = 1
projection_code(::Armon.EulerProjection) = 0; projection_code(::Armon.Euler2ndProjection) = 1
splitting_code(::Armon.SequentialSplitting) = 0; splitting_code(::Armon.GodunovSplitting) = 1
splitting_code(::Armon.StrangSplitting) = 2; splitting_code(::Armon.XOnlySplitting) = 3
splitting_code(::Armon.YOnlySplitting) = 4

function c_test_case(params::ArmonParameters{T}) where {T}
    test = params.test
    tp = Armon.init_test_params(test, T)                       # src/tests.jl:84-121
    sides = (Side.Left, Side.Right, Side.Bottom, Side.Top)
    bc = map(s -> Armon.boundary_condition(test, s), sides)    # src/tests.jl:150-211 -> (u_factor, v_factor)
    CTestCase(test_code(test), test isa Armon.Bizarrium ? 1 : 0,
              tp.high_ρ, tp.low_ρ, tp.high_E, tp.low_E, tp.high_u, tp.low_u, tp.high_v, tp.low_v,
              test isa Armon.Sedov ? test.r : 0.0, Armon.specific_heat_ratio(test),
              map(first, bc), map(last, bc))
end

function c_solver_desc(params::ArmonParameters{T}, state::SolverState) where {T}
    nb = map(s -> Int32(params.neighbours[s] == Armon.MPI.PROC_NULL ? -1 : params.neighbours[s]),
             (Side.Left, Side.Right, Side.Bottom, Side.Top))
    CSolverDesc(CDims(params.N..., params.nghost), params.global_grid..., params.N_origin...,
                Tuple(params.domain_size), Tuple(params.origin),
                riemann_code(state.riemann_scheme), limiter_code(state.riemann_limiter),
                projection_code(state.projection_scheme), splitting_code(state.splitting),
                params.cfl, params.maxtime, params.maxcycle, params.cst_dt, params.Dt, nb,
                params.backend_options.math_mode, 0, params.backend_options.kernel_variant,
                params.backend_options.cuda_graph, c_test_case(params))
end

# ------------------------------------------------------------------------------------------------------------
# Seam 2a: the fused path -- one ccall per solver cycle, the time-step state machine lives on the device
# ------------------------------------------------------------------------------------------------------------
# One block per GPU: ArmonParameters(; use_gpu=true, device=:B200, use_cache_blocking=true,
#     block_size = N_local .+ 2nghost, async_cycle=false, use_threading=false, numa_aware=false, gpu_aware=false)
# (SURVEY.md section 0.10: the only configuration in which the unmodified constructor yields a 1x1 block grid).
the_block(grid::BlockGrid) = only(Armon.device_blocks(grid))

function fused_solver(params::ArmonParameters{T, <:B200Device}, grid::BlockGrid) where {T}
    opts = params.backend_options
    opts.solver != C_NULL && return opts.solver
    blk = the_block(grid); d = Armon.block_device_data(blk); state = blk.state
    desc = Ref(c_solver_desc(params, state)); out = Ref{Ptr{Cvoid}}(C_NULL)
    @b200call(:armon_solver_create, (Ptr{Cvoid}, Ptr{CSolverDesc}, Ptr{Ptr{Cvoid}}), params.device.ctx, desc, out)
    main = [d.ρ.ptr, d.u.ptr, d.v.ptr, d.E.ptr]
    work = [d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr]
    pcg  = [d.p.ptr, d.c.ptr, d.g.ptr]
    @b200call(:armon_solver_bind, (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}),
              out[], main, work, pcg)
    if params.use_MPI && params.proc_size > 1      # NCCL bootstrap: MPI.bcast of the unique id (init_MPI's communicator)
        id = zeros(UInt8, 128)
        params.rank == 0 && @b200call(:armon_comm_unique_id, (Ptr{UInt8},), id)
        Armon.MPI.Bcast!(id, 0, params.cart_comm)
        @b200call(:armon_ctx_comm_init, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), params.device.ctx, id,
                  params.rank, params.proc_size)
    end
    opts.solver = out[]
end

# init_test(params, grid), src/kernels.jl:176-214 (full domain incl. ghosts) + reset!(global_dt)
function Armon.init_test(params::ArmonParameters{<:Any, <:B200Device}, grid::BlockGrid)
    @b200call(:armon_solver_init, (Ptr{Cvoid},), fused_solver(params, grid))
end

# solver_cycle(params, grid), src/solver.jl:288-320: EOS_init + next_time_step + every sweep of split_axes
function Armon.solver_cycle(params::ArmonParameters{<:Any, <:B200Device}, grid::BlockGrid)
    params.compare && return invoke(Armon.solver_cycle, Tuple{ArmonParameters, BlockGrid}, params, grid)  # per-step path
    @b200call(:armon_solver_run, (Ptr{Cvoid}, Int64), fused_solver(params, grid), 1)
    params.backend_options.dirty = true
    return false
end

# next_cycle!(params, global_dt), src/solver_state.jl:145-166: already applied on the device; mirror the scalars
function Armon.next_cycle!(params::ArmonParameters{<:Any, <:B200Device}, global_dt::Armon.GlobalTimeStep)
    st = Ref{CTimeState}()
    @b200call(:armon_solver_state, (Ptr{Cvoid}, Ptr{CTimeState}), params.backend_options.solver, st)
    st[].error == 4 && solver_error(:time, "Invalid time step for cycle $(st[].error_cycle)")
    st[].error == 6 && solver_error(:cpp, "cycle $(st[].error_cycle): the strict mode's IEEE fix-up list overflowed; use math_mode=:ieee")
    global_dt.cycle, global_dt.time = st[].cycle, st[].time
    global_dt.current_dt, global_dt.next_cycle_dt = st[].current_dt, st[].next_cycle_dt
end

# device_to_host!(grid), src/blocking/block_grid.jl:717-729: canonical layout + stale p, c, g first
function Armon.device_to_host!(grid::BlockGrid{<:Any, <:B200Array})
    ensure_canonical(grid.params)
    invoke(Armon.device_to_host!, Tuple{BlockGrid}, grid)
end

# time_loop(params, grid), src/solver.jl:323-403, when nothing needs the host between cycles (silent > 1, no animation):
# the whole `while time < maxtime && cycle < maxcycle` runs on the device.  With silent <= 1 on one rank the per-cycle
# log is produced on the device too (armon_solver_diagnostics) and printed when the loop returns.
function device_time_loop(params::ArmonParameters{<:Any, <:B200Device}, grid::BlockGrid)
    solver = fused_solver(params, grid)
    verbose = params.silent <= 1 && !(params.use_MPI && params.proc_size > 1)
    verbose && @b200call(:armon_solver_diagnostics, (Ptr{Cvoid}, Int32), solver, min(params.maxcycle + 1, 1 << 20))
    @b200call(:armon_solver_time_loop, (Ptr{Cvoid},), solver)
    params.backend_options.dirty = true
    if verbose
        lines = Vector{CCycleDiag}(undef, 4096); n = Ref{Int64}(0)
        while true
            @b200call(:armon_solver_read_diagnostics, (Ptr{Cvoid}, Ptr{CCycleDiag}, Int64, Ptr{Int64}), solver, lines, length(lines), n)
            n[] == 0 && break
            for l in view(lines, 1:n[])
                ΔM = abs(params.initial_mass - l.mass) / params.initial_mass * 100
                ΔE = abs(params.initial_energy - l.energy) / params.initial_energy * 100
                Armon.@printf("Cycle %4d: dt = %.18f, t = %.18f, |ΔM| = %#8.6g%%, |ΔE| = %#8.6g%%\n", l.cycle, l.dt, l.time, ΔM, ΔE)
            end
        end
        @b200call(:armon_solver_diagnostics, (Ptr{Cvoid}, Int32), solver, 0)
    end
end

# Layout of the state between the sweeps of the fused loop (include/armon_b200.h: armon_solver_tiled): 1 band-tiled,
# 0 row-major / transposed pair, -1 not decided yet (before the first cycle).  Informational: the arrays Julia sees are
# canonical whenever `ensure_canonical` has run; ENV["ARMON_B200_TILED"] = "0" keeps the row-major layouts.
function fused_layout_is_tiled(params::ArmonParameters{<:Any, <:B200Device}, grid::BlockGrid)
    tiled = Ref{Int32}(-1)
    @b200call(:armon_solver_tiled, (Ptr{Cvoid}, Ptr{Int32}), fused_solver(params, grid), tiled)
    tiled[]
end

# Kernel of the bit-exact mode (armon_solver_strict_chains): 1 = strict arithmetic on the fast kernel's four-chain
# schedule (default for math_mode strict), 0 = register-prefetch kernel (ENV["ARMON_B200_STRICT"] = "single") or another mode.
function strict_kernel_is_chains(params::ArmonParameters{<:Any, <:B200Device}, grid::BlockGrid)
    chains = Ref{Int32}(0)
    @b200call(:armon_solver_strict_chains, (Ptr{Cvoid}, Ptr{Int32}), fused_solver(params, grid), chains)
    chains[]
end

# ------------------------------------------------------------------------------------------------------------
# Seam 2b: per-step overloads (debug / `compare=true` checkpoints, src/io.jl:185-227): one ccall per
# `@generic_kernel`, same per-block wrapper signatures as the reference (src/kernels.jl:151-230,
# src/riemann_schemes.jl:46-123, src/projection_schemes.jl:44-145, src/halo_exchange.jl:32-36, src/reductions.jl:65-110,
# 271-298).  They are what `solver_cycle` falls back to when `params.compare` is set (see above).
# ------------------------------------------------------------------------------------------------------------
const B200Params = ArmonParameters{<:Any, <:B200Device}
const PF = Ptr{Float64}

# block_domain_range(bsize, steps_range) -> inclusive rectangle in 1-based real-cell coordinates (armon_domain)
function c_domain(blk::LocalTaskBlock, range)
    (nx, ny) = Armon.real_block_size(blk.size)
    ((blx, bly), (trx, try_)) = (Tuple(range[1]), Tuple(range[2]))     # corner offsets of a StepsRanges entry
    CDomain(1 + blx, nx + trx, 1 + bly, ny + try_)
end
c_dims(blk::LocalTaskBlock) = CDims(Armon.real_block_size(blk.size)..., Armon.ghosts(blk.size))
c_axis(state::SolverState) = Cint(Int(state.axis) - 1)
swept_velocity(state::SolverState, d) = state.axis == Axis.X ? d.u : d.v

function Armon.update_EOS!(params::B200Params, state::SolverState, blk::LocalTaskBlock, tc::Armon.TestCase)
    d = Armon.block_device_data(blk)
    @b200call(:armon_perfect_gas_EOS, (Ptr{Cvoid}, CDims, CDomain, Float64, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.EOS),
              Armon.specific_heat_ratio(tc), d.ρ, d.E, d.u, d.v, d.p, d.c, d.g)
end

function Armon.update_EOS!(params::B200Params, state::SolverState, blk::LocalTaskBlock, ::Armon.Bizarrium)
    d = Armon.block_device_data(blk)
    @b200call(:armon_bizarrium_EOS, (Ptr{Cvoid}, CDims, CDomain, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.EOS), d.ρ, d.u, d.v, d.E, d.p, d.c, d.g)
end

function Armon.boundary_conditions!(params::B200Params, state::SolverState, blk::LocalTaskBlock, side::Side.T)
    d = Armon.block_device_data(blk)
    (u_factor, v_factor) = Armon.boundary_condition(state.test_case, side)
    @b200call(:armon_boundary_conditions, (Ptr{Cvoid}, CDims, Cint, Float64, Float64, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), Cint(Int(side) - 1), u_factor, v_factor, d.ρ, d.u, d.v, d.p, d.c, d.g, d.E)
end

function Armon.numerical_fluxes!(params::B200Params, state::SolverState, blk::LocalTaskBlock, ::Armon.RiemannGodunov)
    d = Armon.block_device_data(blk)
    @b200call(:armon_acoustic, (Ptr{Cvoid}, CDims, CDomain, Cint, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.fluxes), c_axis(state),
              d.uˢ, d.pˢ, d.ρ, swept_velocity(state, d), d.p, d.c)
end

function Armon.numerical_fluxes!(params::B200Params, state::SolverState, blk::LocalTaskBlock, ::Armon.RiemannGAD)
    d = Armon.block_device_data(blk)
    @b200call(:armon_acoustic_GAD, (Ptr{Cvoid}, CDims, CDomain, Cint, Float64, Float64, Cint, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.fluxes), c_axis(state),
              state.dt, state.dx, limiter_code(state.riemann_limiter), d.uˢ, d.pˢ, d.ρ, swept_velocity(state, d), d.p, d.c)
end

function Armon.cell_update!(params::B200Params, state::SolverState, blk::LocalTaskBlock)
    d = Armon.block_device_data(blk)
    @b200call(:armon_cell_update, (Ptr{Cvoid}, CDims, CDomain, Cint, Float64, Float64, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.cell_update), c_axis(state),
              state.dx, state.dt, d.uˢ, d.pˢ, d.ρ, swept_velocity(state, d), d.E)
end

function Armon.advection_fluxes!(params::B200Params, state::SolverState, blk::LocalTaskBlock, ::Armon.EulerProjection)
    d = Armon.block_device_data(blk)
    @b200call(:armon_advection_first_order, (Ptr{Cvoid}, CDims, CDomain, Cint, Float64, PF, PF, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.advection), c_axis(state), state.dt,
              d.uˢ, d.ρ, d.u, d.v, d.E, d.work_1, d.work_2, d.work_3, d.work_4)
end

function Armon.advection_fluxes!(params::B200Params, state::SolverState, blk::LocalTaskBlock, ::Armon.Euler2ndProjection)
    d = Armon.block_device_data(blk)
    @b200call(:armon_advection_second_order,
              (Ptr{Cvoid}, CDims, CDomain, Cint, Float64, Float64, PF, PF, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.advection), c_axis(state), state.dx, state.dt,
              d.uˢ, d.ρ, d.u, d.v, d.E, d.work_1, d.work_2, d.work_3, d.work_4)
end

function Armon.euler_projection!(params::B200Params, state::SolverState, blk::LocalTaskBlock)
    d = Armon.block_device_data(blk)
    @b200call(:armon_euler_projection, (Ptr{Cvoid}, CDims, CDomain, Cint, Float64, Float64, PF, PF, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), c_domain(blk, state.steps_ranges.projection), c_axis(state), state.dx, state.dt,
              d.uˢ, d.ρ, d.u, d.v, d.E, d.work_1, d.work_2, d.work_3, d.work_4)
end

# dtCFL_kernel(params, state, blk, ΔX), src/reductions.jl:65-88: returns the host value like `mapreduce`
function Armon.dtCFL_kernel(params::B200Params, ::SolverState, blk::LocalTaskBlock, ΔX)
    ensure_canonical(params)
    d = Armon.block_device_data(blk)
    res = Ref{Float64}(Inf)
    @b200call(:armon_dtCFL, (Ptr{Cvoid}, CDims, PF, PF, PF, Float64, Float64, Ptr{Float64}),
              params.device.ctx, c_dims(blk), d.u, d.v, d.c, ΔX[1], ΔX[2], res)
    return res[]
end

# conservation_vars(params, blk), src/reductions.jl:271-298 -> (mass, energy) of the block
function Armon.conservation_vars(params::B200Params, blk::LocalTaskBlock)
    ensure_canonical(params)
    d = Armon.block_device_data(blk)
    ds = prod(params.domain_size ./ params.global_grid)
    mass, energy = Ref{Float64}(0), Ref{Float64}(0)
    @b200call(:armon_conservation_vars, (Ptr{Cvoid}, CDims, PF, PF, Float64, Ptr{Float64}, Ptr{Float64}),
              params.device.ctx, c_dims(blk), d.ρ, d.E, ds, mass, energy)
    return (mass[], energy[])
end

# init_test(params, blk), src/kernels.jl:176-214: full domain including ghosts (per-step path; the fused path uses
# armon_solver_init above)
function Armon.init_test(params::B200Params, blk::LocalTaskBlock)
    d = Armon.block_device_data(blk)
    tc = Ref(c_test_case(params))
    ds, org = collect(Float64, params.domain_size), collect(Float64, params.origin)
    @b200call(:armon_init_test,
              (Ptr{Cvoid}, CDims, Int64, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{CTestCase},
               PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF),
              params.device.ctx, c_dims(blk), params.N_origin..., params.global_grid..., ds, org, tc,
              d.x, d.y, d.mask, d.ρ, d.E, d.u, d.v, d.p, d.c, d.g, d.uˢ, d.pˢ, d.work_1, d.work_2, d.work_3, d.work_4)
end

end # module
