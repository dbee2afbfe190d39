# B200Meshing.jl -- Julia host shim: the reference's `isosurface` signature on top of libb200iso.so.
#
# Drop-in for the two hot methods of Meshing.jl v0.7.0
#     isosurface(sdf::AbstractArray{T,3}, method::MarchingCubes,      X=-1:1, Y=-1:1, Z=-1:1)   src/marching_cubes.jl:27
#     isosurface(sdf::AbstractArray{T,3}, method::MarchingTetrahedra, X=-1:1, Y=-1:1, Z=-1:1)   src/marching_tetrahedra.jl:129
# and the default forwarder isosurface(A, args...) (src/isosurface.jl:30-32), with the same exported names
# (src/Meshing.jl:9-11), the same @kwdef method structs (src/algorithmtypes.jl:23-40), the same return types
# (Vector{NTuple{3,float(FT)}}, Vector{NTuple{3,Int}}) and the same vertex/face order.
#
# NOTE: Julia is not installed in the build image, so this file could not be executed there; every call it
# makes is mirrored one-to-one by the Python host (meshing.jl_b200/api.py + capi.py), which IS tested against
# the library on a B200.  Layout facts relied upon: Vector{NTuple{3,Float32}} == float[3n],
# Vector{NTuple{3,Float64}} == double[3n], Vector{NTuple{3,Int}} == int64_t[3n] (isbits tuples are stored inline).
#
# There is no CPU fallback: anything but a dense Array{Float32,3} / Array{Float64,3} raises ArgumentError.
module B200Meshing

export isosurface, MarchingCubes, MarchingTetrahedra
# (B200Meshing.isosurface_two_phase is the count -> allocate -> generate form; not exported, the reference has no such name)

const libb200iso = get(ENV, "B200ISO_LIB", joinpath(@__DIR__, "..", "lib", "libb200iso.so"))

abstract type AbstractMeshingAlgorithm end
Base.@kwdef struct MarchingCubes{T} <: AbstractMeshingAlgorithm
    iso::T = 0.0
end
Base.@kwdef struct MarchingTetrahedra{T,E} <: AbstractMeshingAlgorithm
    iso::T = 0.0
    eps::E = 1e-3
end

# struct b200iso_params (include/b200iso.h)
struct Params
    algo::Int32
    iso_is_f32::Int32
    eps_is_f32::Int32
    range_kind::Int32
    iso::Float64
    eps::Float64
    x0::Float64; x1::Float64
    y0::Float64; y1::Float64
    z0::Float64; z1::Float64
    x_offset::Int64
    nx_global::Int64
    field_is_f64::Int32
    x_ghost::Int32
end

const B200ISO_MC, B200ISO_MT = Int32(0), Int32(1)
const B200ISO_HOST, B200ISO_DEVICE = Cint(0), Cint(1)
const B200ISO_ECAPACITY = Cint(-5)

last_error() = unsafe_string(ccall((:b200iso_last_error, libb200iso), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("b200iso error $rc: $(last_error())")

# A handle is not thread-safe and a call is a multi-step conversation with it (extract -> maybe fetch), so handles are
# checked out of a pool for the duration of ONE isosurface call and returned afterwards.  (Keying them by
# Threads.threadid() is not safe: a Julia task may migrate between threads in the middle of a call.)
const handle_pool = Ptr{Cvoid}[]
const handle_pool_lock = ReentrantLock()
function new_handle()
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b200iso_create, libb200iso), Cint, (Ref{Ptr{Cvoid}}, Cint), h, parse(Cint, get(ENV, "B200ISO_DEVICE", "0"))))
    h[]
end
function with_handle(f)
    h = lock(handle_pool_lock) do
        isempty(handle_pool) ? C_NULL : pop!(handle_pool)
    end
    h == C_NULL && (h = new_handle())
    try
        return f(h)
    finally
        lock(handle_pool_lock) do
            push!(handle_pool, h)
        end
    end
end

# typeof(iso)/typeof(eps): Float32 stays Float32; Float64 stays Float64; an Integer behaves like Float32 next to
# a Float32 field (promote_type(Int, Float32) == Float32) and must be exact in Float32.
scalar_kind(v::Float32) = (Float64(v), Int32(1))
scalar_kind(v::Float64) = (v, Int32(0))
function scalar_kind(v::Integer)
    Float32(v) == v || throw(ArgumentError("integer level $v is not exact in Float32"))
    (Float64(v), Int32(1))
end
scalar_kind(v) = throw(ArgumentError("unsupported level type $(typeof(v)) on the B200 path"))

range_kind(::Type{<:Integer}) = Int32(0)
range_kind(::Type{Float32}) = Int32(1)
range_kind(::Type{Float64}) = Int32(2)
range_kind(T) = throw(ArgumentError("unsupported range element type $T on the B200 path"))

# The vertex type follows eltype(first(X)) only (src/marching_cubes.jl:31), the coordinates follow
# LinRange(first(X), last(X), n), i.e. the promotion of both endpoints (:36-38):
#   first is Float64 -> 2; else promote(first, last) is Float32 -> 1; else (Int/Int, Int/Float64, Float32/Float64) -> 0
function axis_kind(R)
    a, b = range_kind(typeof(first(R))), range_kind(typeof(last(R)))
    a == 2 ? Int32(2) : ((a != 2 && b != 2 && (a == 1 || b == 1)) ? Int32(1) : Int32(0))
end

function params(method, X, Y, Z, field_is_f64::Bool)
    kx, ky, kz = axis_kind(X), axis_kind(Y), axis_kind(Z)
    kx == ky == kz || throw(ArgumentError("X, Y, Z must share an element type on the B200 path"))
    iso, isf = scalar_kind(method.iso)
    if method isa MarchingCubes
        algo, eps, epf = B200ISO_MC, 1e-3, Int32(1)
    else
        algo = B200ISO_MT
        eps, epf = scalar_kind(method.eps)
    end
    Params(algo, isf, epf, kx, iso, eps, first(X), last(X), first(Y), last(Y), first(Z), last(Z), 0, 0, Int32(field_is_f64), Int32(0))
end

vertex_eltype(p::Params) =
    (p.field_is_f64 != 0 || p.iso_is_f32 == 0 || p.range_kind == 2 || (p.algo == B200ISO_MT && p.eps_is_f32 == 0)) ? Float64 : Float32

# Capacity guess for the one-shot call: the totals of the previous call with the same shape and method plus 1/8, else
# an estimate from the surface area an isosurface of this grid typically has (the 1024^3 gyroid: 38.7 n^2 MC vertices,
# 19.3 n^2 faces; MT 27.7 n^2 / 55.2 n^2).  Over-allocation is free: untouched pages of an `undef` Vector are never
# backed by memory, and `resize!` to a smaller length does not copy.
const capacity_memo = Dict{Tuple{NTuple{3,Int},DataType,Int32},NTuple{2,Int}}()
const capacity_memo_lock = ReentrantLock()
function capacity_guess(key, dims, algo)
    m = lock(() -> get(capacity_memo, key, nothing), capacity_memo_lock)
    m === nothing || return (m[1] + m[1] ÷ 8 + 1024, m[2] + m[2] ÷ 8 + 1024)
    nx, ny, nz = max.(dims .- 1, 0)
    area = (nx * ny + ny * nz + nx * nz) / 3
    nvox = nx * ny * nz
    algo == B200ISO_MC ? (Int(floor(min(64area, 12nvox))) + 1024, Int(floor(min(32area, 5nvox))) + 1024) :
                         (Int(floor(min(48area, 7nvox))) + 1024, Int(floor(min(96area, 12nvox))) + 1024)
end

# The drop-in body: ONE b200iso_extract_host into arrays sized by the guess -- the mesh streams out of the GPU while the
# field still streams in (x-slab pipeline) -- then trim.  If the guess was short (B200ISO_ECAPACITY, exact totals
# returned) the mesh is fetched into exact arrays from the slabs still resident on the device
# (b200iso_extract_host_resident): the field is uploaded once in every case.
function _isosurface(sdf::Union{Array{Float32,3},Array{Float64,3}}, method, X, Y, Z)
    nx, ny, nz = size(sdf)
    p = Ref(params(method, X, Y, Z, eltype(sdf) === Float64))
    VT = vertex_eltype(p[])
    key = (size(sdf), eltype(sdf), p[].algo)
    vcap, fcap = capacity_guess(key, size(sdf), p[].algo)
    nv, nf, f64 = Ref{Int64}(0), Ref{Int64}(0), Ref{Cint}(0)
    vts = Vector{NTuple{3,VT}}(undef, vcap)
    fcs = Vector{NTuple{3,Int}}(undef, fcap)
    with_handle() do h
        rc = GC.@preserve sdf vts fcs ccall((:b200iso_extract_host, libb200iso), Cint,
            (Ptr{Cvoid}, Ref{Params}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Int64}, Int64,
             Ref{Int64}, Ref{Int64}, Ref{Cint}),
            h, p, pointer(sdf), nx, ny, nz, nx, pointer(vts), vcap, pointer(fcs), fcap, nv, nf, f64)
        if rc == B200ISO_ECAPACITY
            vts = Vector{NTuple{3,VT}}(undef, nv[])
            fcs = Vector{NTuple{3,Int}}(undef, nf[])
            GC.@preserve vts fcs check(ccall((:b200iso_extract_host_resident, libb200iso), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Int64, Ref{Int64}, Ref{Int64}),
                h, pointer(vts), nv[], pointer(fcs), nf[], nv, nf))
        else
            check(rc)
        end
    end
    @assert (f64[] != 0) == (VT === Float64)
    lock(() -> (capacity_memo[key] = (Int(nv[]), Int(nf[]))), capacity_memo_lock)
    resize!(vts, nv[]), resize!(fcs, nf[])
end

# The two-phase pair (count: the caller learns the sizes; generate: into exact arrays) for callers that cannot
# over-allocate; no overlap between the upload and the download.
function isosurface_two_phase(sdf::Union{Array{Float32,3},Array{Float64,3}}, method::Union{MarchingCubes,MarchingTetrahedra},
                              X=-1:1, Y=-1:1, Z=-1:1)
    nx, ny, nz = size(sdf)
    p = Ref(params(method, X, Y, Z, eltype(sdf) === Float64))
    VT = vertex_eltype(p[])
    nv, nf, f64 = Ref{Int64}(0), Ref{Int64}(0), Ref{Cint}(0)
    with_handle() do h
        GC.@preserve sdf check(ccall((:b200iso_count, libb200iso), Cint,
            (Ptr{Cvoid}, Ref{Params}, Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Ref{Int64}, Ref{Int64}, Ref{Cint}),
            h, p, pointer(sdf), B200ISO_HOST, nx, ny, nz, nx, nv, nf, f64))
        vts = Vector{NTuple{3,VT}}(undef, nv[])
        fcs = Vector{NTuple{3,Int}}(undef, nf[])
        GC.@preserve vts fcs check(ccall((:b200iso_generate, libb200iso), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Cint, Int64),
            h, pointer(vts), pointer(fcs), B200ISO_HOST, 0))
        vts, fcs
    end
end

# same positional signature and defaults as the reference
function isosurface(sdf::AbstractArray{T,3}, method::Union{MarchingCubes,MarchingTetrahedra}, X=-1:1, Y=-1:1, Z=-1:1) where {T}
    (sdf isa Array{Float32,3} || sdf isa Array{Float64,3}) ||
        throw(ArgumentError("the B200 path accepts a dense Array{Float32,3} or Array{Float64,3} (got $(typeof(sdf))); there is no CPU fallback"))
    _isosurface(sdf, method, X, Y, Z)
end
# src/isosurface.jl:30-32
isosurface(A::AbstractArray{T,3}, args...) where {T} = isosurface(A, MarchingCubes(), args...)

end # module
