# JustPICSM100aExt.jl -- reference-side binding of libjustpic_sm100a.so.
#
# What a JustPIC.jl maintainer would add (as a package extension next to
# ext/JustPICCUDAExt.jl) to route the particle-in-cell hot path of
# `Particles{CUDABackend}` through the B200-native library instead of the
# KernelAbstractions kernels.  It defines more specific methods of the L4
# launchers (the functions that today end in `launch!(ka_backend(x), kernel!, ...)`,
# src/launch.jl:65-69) and `ccall`s the C ABI of include/justpic_c.h.
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia.  It is kept thin
# on purpose -- pointer plumbing only, every decision lives behind the C ABI, which
# the Python mirror (justpic/jl_b200/api.py) exercises call for call.
#
# Memory: CuCellArray{T,N,0} (blocklength 0, ext/JustPICCUDAExt.jl:26-30) stores
# `data[C, S, 1]` column-major, i.e. element (cell c, slot s) at c + s*C -- exactly
# the layout the library expects, so `pointer(A.data)` is passed as is; nothing is
# copied or converted, and `Array(particles)`, `copy`, JLD2 checkpointing keep working.
module JustPICSM100aExt

using CUDA, JustPIC, CellArrays
import MPI
using CUDA: CUDABackend
import JustPIC: Particles, Euler, RungeKutta2, RungeKutta4, AbstractAdvectionIntegrator

const libjustpic = get(ENV, "JUSTPIC_SM100A_LIB", "libjustpic_sm100a.so")

# ---- C structs (include/justpic_c.h) ---------------------------------------
struct JpGridDesc
    ndim::Int32
    n::NTuple{3, Int32}
    S::Int32
    uniform::Int32
    xv::NTuple{3, Ptr{Float64}}
    xc::NTuple{3, Ptr{Float64}}
    xvel::NTuple{9, Ptr{Float64}}      # [comp][dim], row-major like the C array
    nvel::NTuple{9, Int32}
end

struct JpParticles
    coords::NTuple{3, CuPtr{Float64}}
    index::CuPtr{UInt8}
end

check(rc::Cint, who) = rc == 0 ? nothing :
    (msg = unsafe_string(ccall((:jp_last_error, libjustpic), Cstring, ()));
     rc == -1 ? throw(ArgumentError("$who: $msg")) : error("$who failed ($rc): $msg"))

# ---- one context per Particles object (created lazily, cached by objectid) --
const CONTEXTS = Dict{UInt, Ptr{Cvoid}}()
const HOST_GRIDS = Dict{UInt, Any}()       # keeps the host copies of the grid vectors alive

pad3(t::NTuple{2}, x) = (t[1], t[2], x)
pad3(t::NTuple{3}, x) = t

function context(p::Particles{CUDABackend, N}) where {N}
    get!(CONTEXTS, objectid(p)) do
        # range grids -> scalar spacings (particles_utils.jl:137-140), arrays -> diff() (:76-79)
        uniform = p.di.vertex[1] isa Number
        host(x) = collect(Float64, Array(x))
        xv = map(host, p.xvi); xc = map(host, p.xci)
        xvel = ntuple(c -> map(host, p.xi_vel[c]), Val(N))
        HOST_GRIDS[objectid(p)] = (xv, xc, xvel)
        nullp = Ptr{Float64}(0)
        xvel9 = ntuple(i -> (c = (i - 1) ÷ 3 + 1; d = (i - 1) % 3 + 1; (c <= N && d <= N) ? pointer(xvel[c][d]) : nullp), 9)
        nvel9 = ntuple(i -> (c = (i - 1) ÷ 3 + 1; d = (i - 1) % 3 + 1; (c <= N && d <= N) ? Int32(length(xvel[c][d])) : Int32(0)), 9)
        desc = Ref(JpGridDesc(Int32(N), pad3(Int32.(size(p.index)), Int32(1)), Int32(p.max_xcell), Int32(uniform),
                              pad3(map(pointer, xv), nullp), pad3(map(pointer, xc), nullp), xvel9, nvel9))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:jp_ctx_create, libjustpic), Cint, (Ref{JpGridDesc}, Cint, Ref{Ptr{Cvoid}}),
                    desc, CUDA.deviceid(CUDA.device()), out), "jp_ctx_create")
        out[]
    end
end

cptr(A::CellArray) = pointer(A.data)
jp(p::Particles{CUDABackend, N}) where {N} =
    Ref(JpParticles(pad3(map(cptr, p.coords), CuPtr{Float64}(0)), cptr(p.index)))
stream() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
argptrs(args) = CuPtr{Float64}[cptr(a) for a in args]
# the reference's launch! synchronises after every kernel (src/launch.jl:60-69); switch off for stream-ordered time loops and
# while capturing a CUDA graph (every library call is asynchronous on the task's stream)
const SYNC_EACH_CALL = Ref(true)
done() = SYNC_EACH_CALL[] && CUDA.synchronize()

scheme(::Euler) = (Int32(0), 0.0)
scheme(m::RungeKutta2) = (Int32(1), Float64(m.α))
scheme(::RungeKutta4) = (Int32(2), 0.0)

# ---- L4 launchers ------------------------------------------------------------
# advection!(particles, method, V, grid_vi, dt, dxi)   src/Particles/Advection/advection.jl:35-62
function JustPIC.advection!(p::Particles{CUDABackend, N}, method::AbstractAdvectionIntegrator, V, grid_vi::NTuple{N, NTuple{N, T}}, dt, dxi) where {N, T}
    s, α = scheme(method)
    Vp = CuPtr{Float64}[pointer(v) for v in V]
    check(ccall((:jp_advect, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, Int32, Float64, Ptr{CuPtr{Float64}}, Float64, Ptr{Cvoid}),
                context(p), jp(p), s, α, Vp, Float64(dt), stream()), "advection!")
    done()
end

# advection! split for the halo overlap (jp_advect_region): region 1 = bricks holding one of the two outermost cell layers,
# region 2 = the rest.  A multi-GPU time loop (scripts/temperature_advection3D_MPI.jl:83-91) becomes
#     advection_region!(p, method, V, dt, 1); ev = CUDA.CuEvent(); CUDA.record(ev)
#     CUDA.stream!(side_stream) do; CUDA.wait(ev); update_cell_halo!(p.coords..., args..., p.index); CUDA.record(done_ev); end
#     advection_region!(p, method, V, dt, 2); CUDA.wait(done_ev)
#     move_particles!(p, args)
function advection_region!(p::Particles{CUDABackend}, method::AbstractAdvectionIntegrator, V, dt, region::Integer)
    s, α = scheme(method)
    Vp = CuPtr{Float64}[pointer(v) for v in V]
    check(ccall((:jp_advect_region, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, Int32, Float64, Ptr{CuPtr{Float64}}, Float64, Int32, Ptr{Cvoid}),
                context(p), jp(p), s, α, Vp, Float64(dt), Int32(region), stream()), "advection!")
    done()
end

# move_particles!(particles, grid, args, dxi)           src/Particles/move_safe.jl:23-49
function JustPIC.move_particles!(p::Particles{CUDABackend}, grid::NTuple{N}, args, dxi) where {N}
    a = argptrs(args)
    check(ccall((:jp_move, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, Ptr{CuPtr{Float64}}, Int32, Ptr{Cvoid}),
                context(p), jp(p), a, Int32(length(a)), stream()), "move_particles!")
    done()
end

# inject_particles!(particles, args, grid, di)          src/Particles/injection.jl:21-53
const INJECT_STEP = Ref{UInt32}(0)
function JustPIC.inject_particles!(p::Particles{CUDABackend}, args, grid::NTuple{N}, di) where {N}
    a = argptrs(args)
    check(ccall((:jp_inject, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, Ptr{CuPtr{Float64}}, Int32, Int32, UInt64, UInt32, Ptr{Cvoid}),
                context(p), jp(p), a, Int32(length(a)), Int32(p.min_xcell), UInt64(42), INJECT_STEP[], stream()), "inject_particles!")
    INJECT_STEP[] += UInt32(1)
    done()
end

# force_injection!(particles, p_new, fields, values)     src/Particles/forced_injection.jl:16-29
# p_new is an array of user points (anything with getindex / isnan): repacked once into N CellArray-shaped
# coordinate arrays, NaN in component 1 marking an inactive point, which is what the library tests per cell.
function JustPIC.force_injection!(p::Particles{CUDABackend}, p_new, fields::NTuple{NF, Any}, values::NTuple{NF, Any}) where {NF}
    N = length(p.coords)
    host = Array(p_new)
    comps = ntuple(d -> CuArray(Float64[(d == 1 && isnan(q)) ? NaN : Float64(q[d]) for q in host]), N)
    pn = CuPtr{Float64}[pointer(c) for c in comps]
    f = argptrs(fields)
    v = Float64[Float64(x) for x in values]
    GC.@preserve comps check(ccall((:jp_force_injection, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, Ptr{CuPtr{Float64}}, Ptr{CuPtr{Float64}}, Ptr{Float64}, Int32, Ptr{Cvoid}),
                context(p), jp(p), pn, f, v, Int32(length(f)), stream()), "force_injection!")
    done()
end

# clean_particles!(particles, grid, args)               src/Particles/move_safe.jl:289-295
function JustPIC.clean_particles!(p::Particles{CUDABackend}, grid, args)
    a = argptrs(args)
    check(ccall((:jp_clean, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, Ptr{CuPtr{Float64}}, Int32, Ptr{Cvoid}),
                context(p), jp(p), a, Int32(length(a)), stream()), "clean_particles!")
    done()
end

# grid2particle!(Fp, xvi, F, particles, di)             src/Interpolations/grid_to_particle.jl:28-35
function JustPIC.grid2particle!(Fp::CellArray, xvi, F::CuArray, p::Particles{CUDABackend}, di)
    check(ccall((:jp_grid2particle, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                context(p), jp(p), cptr(Fp), pointer(F), stream()), "grid2particle!")
    done()
end

# centroid2particle!(Fp, xci, F, particles, di)         src/Interpolations/centroid_to_particle.jl:15-20
function JustPIC.centroid2particle!(Fp::CellArray, xci, F::CuArray, p::Particles{CUDABackend}, di)
    check(ccall((:jp_centroid2particle, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                context(p), jp(p), cptr(Fp), pointer(F), stream()), "centroid2particle!")
    done()
end

# particle2grid!(F, Fp, particles)                      src/Interpolations/particle_to_grid.jl:23-28
function JustPIC.particle2grid!(F::CuArray, Fp::CellArray, p::Particles{CUDABackend})
    check(ccall((:jp_particle2grid, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                context(p), jp(p), pointer(F), cptr(Fp), stream()), "particle2grid!")
    done()
end

# particle2centroid!(F, Fp, xci, particles, di)         src/Interpolations/particle_to_grid_centroid.jl:12-16
function JustPIC.particle2centroid!(F::CuArray, Fp::CellArray, xci::NTuple, p::Particles{CUDABackend}, di)
    check(ccall((:jp_particle2centroid, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                context(p), jp(p), pointer(F), cptr(Fp), stream()), "particle2centroid!")
    done()
end

# phase_ratios_center!(phase_ratios, particles, phases) src/PhaseRatios/centers.jl:3-11
function JustPIC.phase_ratios_center!(pr::JustPIC.PhaseRatios{CUDABackend}, p::Particles{CUDABackend}, phases)
    K = JustPIC.numphases(pr)
    check(ccall((:jp_phase_ratios_center, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Int32, Ptr{Cvoid}),
                context(p), jp(p), cptr(pr.center), cptr(phases), Int32(K), stream()), "phase_ratios_center!")
    done()
end

# ---- section 8 "next" rows ------------------------------------------------------------------------
const FACE_DIM = Dict(:x => 0, :y => 1, :z => 2)
const MID_PLANE = Dict(:xy => 0, :yz => 1, :xz => 2)
const PhaseCall = (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Int32, Ptr{Cvoid})
const PhaseCallDim = (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, Int32, Int32, Ptr{Cvoid})

# phase_ratios_vertex!(phase_ratios, particles, phases)            src/PhaseRatios/vertices.jl:4-13
function JustPIC.phase_ratios_vertex!(pr::JustPIC.PhaseRatios{CUDABackend}, p::Particles{CUDABackend}, phases)
    check(ccall((:jp_phase_ratios_vertex, libjustpic), Cint, PhaseCall,
                context(p), jp(p), cptr(pr.vertex), cptr(phases), Int32(JustPIC.numphases(pr)), stream()), "phase_ratios_vertex!")
    done()
end

# phase_ratios_face!(phase_face, particles, phases, dimension)      src/PhaseRatios/midpoints.jl:3-24
function JustPIC.phase_ratios_face!(face::CellArray, p::Particles{CUDABackend}, phases, dimension::Symbol)
    haskey(FACE_DIM, dimension) || throw(ArgumentError("dimension must be :x, :y or :z"))
    check(ccall((:jp_phase_ratios_face, libjustpic), Cint, PhaseCallDim,
                context(p), jp(p), cptr(face), cptr(phases), Int32(JustPIC.nphases(face)), Int32(FACE_DIM[dimension]), stream()),
          "phase_ratios_face!")
    done()
end

# phase_ratios_midpoint!(phase_midpoint, particles, phases, dimension)   src/PhaseRatios/midpoints.jl:115-126
function JustPIC.phase_ratios_midpoint!(mid::CellArray, p::Particles{CUDABackend}, phases, dimension::Symbol)
    haskey(MID_PLANE, dimension) || throw("Unknown dimensions. Valid dimensions are :xy, :yz, :xz")
    check(ccall((:jp_phase_ratios_midpoint, libjustpic), Cint, PhaseCallDim,
                context(p), jp(p), cptr(mid), cptr(phases), Int32(JustPIC.nphases(mid)), Int32(MID_PLANE[dimension]), stream()),
          "phase_ratios_midpoint!")
    done()
end
# update_phase_ratios!(phase_ratios, particles, phases)            src/PhaseRatios/utils.jl:15-41
# one call for all outputs; JUSTPIC_PHASE_MODE = 0 literal kernels (bit-exact), 1 fused one-pass mode (1e-12, default)
const PHASE_MODE = Ref{Int32}(parse(Int32, get(ENV, "JUSTPIC_PHASE_MODE", "1")))
function JustPIC.update_phase_ratios!(pr::JustPIC.PhaseRatios{CUDABackend}, p::Particles{CUDABackend, N}, phases) where {N}
    faces = CuPtr{Float64}[cptr(pr.Vx), cptr(pr.Vy)]
    N == 3 && push!(faces, cptr(pr.Vz))
    mids = N == 3 ? CuPtr{Float64}[cptr(pr.xy), cptr(pr.yz), cptr(pr.xz)] : CuPtr{Float64}[]
    check(ccall((:jp_update_phase_ratios, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, Int32, CuPtr{Float64}, CuPtr{Float64}, Ptr{CuPtr{Float64}}, Ptr{CuPtr{Float64}}, Int32, Ptr{Cvoid}),
                context(p), jp(p), cptr(phases), Int32(JustPIC.numphases(pr)), cptr(pr.center), cptr(pr.vertex), faces, mids, PHASE_MODE[], stream()),
          "update_phase_ratios!")
    done()
end

# inject_particles_phase!(particles, particles_phases, args, fields, grid, grid_center, di, di_center)   src/Particles/injection.jl:153-200
function JustPIC.inject_particles_phase!(p::Particles{CUDABackend, N}, phases, args, fields, grid::NTuple{N}, grid_center, di, di_center) where {N}
    seed, step = UInt64(42), INJECT_STEP[]
    INJECT_STEP[] += UInt32(1)
    ptrs = argptrs(args)
    fptrs = CuPtr{Float64}[pointer(f) for f in fields]
    kinds = Int32[size(f) == size(p.index) ? 1 : 0 for f in fields]        # centre field iff one value per cell (:295)
    check(ccall((:jp_inject_phase, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, Ptr{CuPtr{Float64}}, Ptr{CuPtr{Float64}}, Ptr{Int32}, Int32, Int32, UInt64, UInt32, Ptr{Cvoid}),
                context(p), jp(p), cptr(phases), ptrs, fptrs, kinds, Int32(length(ptrs)), Int32(p.min_xcell), seed, UInt32(step), stream()),
          "inject_particles_phase!")
    done()
end

# advection_LinP! / advection_MQS!(particles, method, V, grid_vi, dt, dxi)   advection_LinP.jl:28-60, advection_MQS.jl:31-60
for (fn, interp) in ((:advection_LinP!, 1), (:advection_MQS!, 2))
    @eval function JustPIC.$fn(p::Particles{CUDABackend, N}, method::AbstractAdvectionIntegrator, V, grid_vi::NTuple{N, NTuple{N}}, dt, dxi) where {N}
        s, α = scheme(method)
        Vp = CuPtr{Float64}[pointer(v) for v in V]
        check(ccall((:jp_advect_interp, libjustpic), Cint,
                    (Ptr{Cvoid}, Ref{JpParticles}, Int32, Float64, Ptr{CuPtr{Float64}}, Float64, Int32, Ptr{Cvoid}),
                    context(p), jp(p), s, α, Vp, Float64(dt), Int32($interp), stream()), $(string(fn)))
        done()
    end
end

# grid2particle_flip!(Fp, xvi, F, F0, particles; α)               src/Interpolations/grid_to_particle.jl:125-138
function JustPIC.grid2particle_flip!(Fp::CellArray, xvi, F::CuArray, F0::CuArray, p::Particles{CUDABackend}; α = 0.0)
    check(ccall((:jp_grid2particle_flip, libjustpic), Cint,
                (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Float64, Ptr{Cvoid}),
                context(p), jp(p), cptr(Fp), pointer(F), pointer(F0), Float64(α), stream()), "grid2particle_flip!")
    done()
end

# subgrid_diffusion!(pT, T_grid, ΔT_grid, subgrid_arrays, particles, dt; d) / subgrid_diffusion_centroid!   src/Physics/subgrid_diffusion.jl:55-113
for (fn, centroid) in ((:subgrid_diffusion!, 0), (:subgrid_diffusion_centroid!, 1))
    @eval function JustPIC.$fn(pT::CellArray, T_grid::CuArray, ΔT_grid::CuArray, sa::JustPIC.SubgridDiffusionCellArrays, p::Particles{CUDABackend}, dt; d = 1.0)
        ext = Int32[size(ΔT_grid)...]
        check(ccall((:jp_subgrid_diffusion, libjustpic), Cint,
                    (Ptr{Cvoid}, Ref{JpParticles}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Int32}, CuPtr{Float64}, CuPtr{Float64},
                     CuPtr{Float64}, CuPtr{Float64}, Float64, Float64, Int32, Ptr{Cvoid}),
                    context(p), jp(p), cptr(pT), pointer(T_grid), pointer(ΔT_grid), ext, cptr(sa.pT0), cptr(sa.pΔT), cptr(sa.dt₀),
                    pointer(sa.ΔT_subgrid), Float64(dt), Float64(d), Int32($centroid), stream()), $(string(fn)))
        done()
    end
end

# init_particles: allocation stays in Julia (cell_array, src/launch.jl:101-108; the
# CuCellArrays are owned by Julia), only the fill kernel (fill_coords_index!,
# particles_utils.jl:168-194) is replaced.  Called from the tail of the reference's
# init_particles instead of `launch!(..., fill_coords_index!, ...)`.
function fill_coords_index!(p::Particles{CUDABackend}; seed::UInt64 = UInt64(42))
    check(ccall((:jp_init_particles, libjustpic), Cint, (Ptr{Cvoid}, Ref{JpParticles}, Int32, UInt64, Ptr{Cvoid}),
                context(p), jp(p), Int32(p.nxcell), seed, stream()), "init_particles")
    done()
end

# update_cell_halo!(particles.coords..., args..., particles.index)      src/CellArrays/ImplicitGlobalGrid.jl:36-41
# ONE library call replaces the per-CellArray ImplicitGlobalGrid.update_halo! calls: pack kernels, the x -> y -> z schedule and
# the NCCL transport (ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd) are inside jp_halo_exchange.  The host supplies an
# NCCL communicator -- NCCL.jl's `comm.handle`, or one made by jp_comm_init from an id broadcast over MPI -- and the neighbour
# ranks of ImplicitGlobalGrid's Cartesian topology (MPI.Cart_shift on IGG's communicator; -1 = no neighbour).
const HALO = Ref{Any}(nothing)          # (comm::Ptr{Cvoid}, nbr::Vector{Int32} of length 6: left / right per dimension)

function init_halo!(mpicomm, cart_comm; device = CUDA.deviceid(CUDA.device()))
    id = zeros(UInt8, 128)
    MPI.Comm_rank(mpicomm) == 0 && check(ccall((:jp_comm_unique_id, libjustpic), Cint, (Ptr{UInt8},), id), "jp_comm_unique_id")
    MPI.Bcast!(id, 0, mpicomm)
    comm = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:jp_comm_init, libjustpic), Cint, (Ptr{UInt8}, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                id, Int32(MPI.Comm_size(mpicomm)), Int32(MPI.Comm_rank(mpicomm)), Int32(device), comm), "jp_comm_init")
    nbr = fill(Int32(-1), 6)
    for d in 0:(MPI.Cartdim_get(cart_comm) - 1)
        l, r = MPI.Cart_shift(cart_comm, d, 1)
        nbr[2d + 1] = l == MPI.PROC_NULL ? Int32(-1) : Int32(l)
        nbr[2d + 2] = r == MPI.PROC_NULL ? Int32(-1) : Int32(r)
    end
    HALO[] = (comm[], nbr)
end

# the reference's signature takes bare CellArrays; the Particles object (for the context) is found through its index array
function update_cell_halo!(p::Particles{CUDABackend}, args::Vararg{CellArray})
    comm, nbr = HALO[]
    arrays = CuPtr{Float64}[cptr(a) for a in (p.coords..., args...)]
    check(ccall((:jp_halo_exchange, libjustpic), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{CuPtr{Float64}}, Int32, CuPtr{UInt8}, Ptr{Cvoid}),
                context(p), comm, nbr, arrays, Int32(length(arrays)), cptr(p.index), stream()), "update_cell_halo!")
    done()
end

# update_halo!(A) of a plain (staggered) grid array on the same decomposition -- the velocity ghost layers when V comes from a solver
function update_halo!(p::Particles{CUDABackend}, A::CuArray{Float64})
    comm, nbr = HALO[]
    ext = Int32[size(A)...]
    check(ccall((:jp_halo_exchange_grid, libjustpic), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, CuPtr{Float64}, Ptr{Int32}, Ptr{Cvoid}),
                context(p), comm, nbr, pointer(A), ext, stream()), "update_halo!")
    done()
end

# dt = min(dx / MPI.Allreduce(maximum(abs.(V)), MPI.MAX, comm)) (scripts/temperature_advection3D_MPI.jl:71) without leaving the GPU
function allreduce_max!(buf::CuArray{Float64})
    check(ccall((:jp_allreduce_max, libjustpic), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Int32, Ptr{Cvoid}),
                HALO[][1], pointer(buf), Int32(length(buf)), stream()), "jp_allreduce_max")
    done(); buf
end

# Array(CA) / Array(T, CA) for a device CellArray                  src/CellArrays/conversion.jl:19-43
# CuArray(CA) / CuArray(T, CA) for a host CellArray                ext/JustPICCUDAExt.jl:166-179
# The reference permutes on the host side of the copy (permutedims(CA.data, (3, 2, 1)) allocates a second full
# array); here the permutation (+ optional Float64 <-> Float32 conversion; Bool stays Bool) is one device
# kernel into a scratch CuArray and the transfer is a single contiguous copy.
const JP_DTYPE = Dict(Float64 => Int32(0), Float32 => Int32(1), Bool => Int32(2))
function permute_layout!(dst::CuArray, src::CuArray, ncells::Integer, ncomp::Integer, to_host::Bool)
    check(ccall((:jp_cellarray_permute, libjustpic), Cint,
                (Ptr{Cvoid}, CuPtr{Cvoid}, Int32, CuPtr{Cvoid}, Int32, Int64, Int32, Int32, Ptr{Cvoid}),
                C_NULL, pointer(src), JP_DTYPE[eltype(src)], pointer(dst), JP_DTYPE[eltype(dst)], Int64(ncells), Int32(ncomp),
                Int32(to_host ? 0 : 1), stream()), "jp_cellarray_permute")
    done()
    return dst
end
function Base.Array(::Type{T}, CA::CellArray{<:Any, <:Any, 0, <:CuArray}) where {T <: Number}      # device -> host image
    ni, S = size(CA), length(eltype(CA))
    Td = eltype(eltype(CA)) === Bool ? Bool : T
    tmp = permute_layout!(CuArray{Td}(undef, 1, S, prod(ni)), CA.data, prod(ni), S, true)
    CA_cpu = JustPIC.CPU_CellArray(SVector{S, Td}, undef, ni)
    copyto!(CA_cpu.data, tmp)
    return CA_cpu
end
function CUDA.CuArray(::Type{T}, CA::CellArray{<:Any, <:Any, 1, <:Array}) where {T <: Number}       # host image -> device
    ni, S = size(CA), length(eltype(CA))
    Td = eltype(eltype(CA)) === Bool ? Bool : T
    CA_gpu = CuCellArray(SVector{S, Td}, undef, ni...)
    permute_layout!(CA_gpu.data, CuArray(CA.data), prod(ni), S, false)
    return CA_gpu
end

# advection -> move hand-off (JP_OPT_ADVECT_CLASSIFY = 5, include/justpic_c.h): advection! also leaves
# move_particles!' classification words; results are bit-identical.  Opt-in because the library cannot see
# writes to coords / index made between the two calls by other code: JUSTPIC_ADVECT_CLASSIFY=1.
const ADVECT_CLASSIFY = Ref{Int32}(parse(Int32, get(ENV, "JUSTPIC_ADVECT_CLASSIFY", "0")))
set_handoff!(p::Particles{CUDABackend}, on::Bool = ADVECT_CLASSIFY[] != 0) =
    check(ccall((:jp_set_option, libjustpic), Cint, (Ptr{Cvoid}, Int32, Int32), context(p), Int32(5), Int32(on)), "jp_set_option")

# move -> interpolation hand-off (JP_OPT_MOVE_INTERP = 7): the last pass of move_particles! also leaves particle2grid!'s cell sums
# of `Fp` and the centre phase ratios of `phases` (both must be among the args of move_particles!); the following
# particle2grid!(F, Fp, particles) / phase_ratios_center!(ratios, particles, phases) then do not read the particles again.
# Bit-identical results; opt-in for the same reason as above.
function set_interp_handoff!(p::Particles{CUDABackend}, Fp::Union{CellArray, Nothing}, phases::Union{CellArray, Nothing}, nphases::Integer = 0; on::Bool = true)
    check(ccall((:jp_set_option, libjustpic), Cint, (Ptr{Cvoid}, Int32, Int32), context(p), Int32(7), Int32(on)), "jp_set_option")
    check(ccall((:jp_move_interp_fields, libjustpic), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Int32),
                context(p), Fp === nothing ? CuPtr{Float64}(0) : cptr(Fp), phases === nothing ? CuPtr{Float64}(0) : cptr(phases), Int32(nphases)),
          "jp_move_interp_fields")
end

# slot policy of move_particles! (JP_OPT_MOVE_POLICY = 4): 0 = the reference's rule (default), 1 = "compact", 2 = "dense" -- the two
# opt-in deviations keep the same particles in the same cells but pack them into lower slots (include/justpic_c.h); JUSTPIC_MOVE_POLICY=0|1|2
set_move_policy!(p::Particles{CUDABackend}, policy::Integer = parse(Int32, get(ENV, "JUSTPIC_MOVE_POLICY", "0"))) =
    check(ccall((:jp_set_option, libjustpic), Cint, (Ptr{Cvoid}, Int32, Int32), context(p), Int32(4), Int32(policy)), "jp_set_option")

# CUDA graphs (JP_OPT_GRAPH_STEP_OFFSET = 10): once a time step has run eagerly (workspaces sized), the same calls can be captured --
#     graph = CUDA.capture() do; advection!(...); move_particles!(...); inject_particles!(...); particle2grid!(...); end
#     exec = CUDA.instantiate(graph); CUDA.launch(exec)            # one launch per step
# -- provided the shim's per-call synchronisation is off while capturing (SYNC_EACH_CALL[] = false).  The library notices the capture
# (cudaStreamIsCapturing), skips move_particles!' asynchronous size feedback and refuses allocations with a message.  Arguments are
# frozen; inject_particles!' step counter advances on the device: replay i injects with step + i.  Reset the offset before going back
# to eager calls.
graph_step_offset!(p::Particles{CUDABackend}, value::Integer = 0) =
    check(ccall((:jp_set_option, libjustpic), Cint, (Ptr{Cvoid}, Int32, Int32), context(p), Int32(10), Int32(value)), "jp_set_option")

# after a write to coords / index / a registered field that did not go through this extension
invalidate_handoffs!(p::Particles{CUDABackend}) = check(ccall((:jp_invalidate_handoffs, libjustpic), Cint, (Ptr{Cvoid},), context(p)), "jp_invalidate_handoffs")

end # module
