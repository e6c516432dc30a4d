# capi.jl — the C ABI of include/particulator_b200.h: library handle, struct mirrors, status handling.

"Path of the shared library; override with ENV[\"PARTICULATOR_B200_LIB\"]."
const LIB = get(ENV, "PARTICULATOR_B200_LIB", "libparticulator_b200")

# ---- struct mirrors (same field order and padding as the header) -------------------------------------------------------
struct ProcessDesc
    kind::Int32
    aux::Int32
    par::NTuple{6,Float64}
end
struct FieldDesc
    kind::Int32
    _pad::Int32
    par::NTuple{7,Float64}
end
struct ForcingDesc
    kind::Int32
    species_mask::UInt32
    e::FieldDesc
    b::FieldDesc
    nel::Float64
    I::Float64
    Tcut::Float64
    cheb_id::Int32
    _pad::Int32
end
struct PusherDesc
    kind::Int32
    restrict_mask::UInt32
    nforcings::Int32
    _pad::Int32
    forcing::NTuple{4,ForcingDesc}
end
struct WallDesc
    species::Int32
    coord::Int32
    v::Float64
    drop::Int32
    _pad::Int32
end
struct CallbackDesc
    nwalls::Int32
    count_collisions::Int32
    wall::NTuple{4,WallDesc}
end
struct DiagOut
    n::Int64
    nactive::Int64
    weight::Float64
    wenergy::Float64
    maxenergy::Float64
    wx::NTuple{3,Float64}
    wx2::NTuple{3,Float64}
    wr2::Float64
end
struct AdvanceStats
    passes::Int64
    substeps::Int64
    rows::Int64
    births::Int64
    launches::Int64
    main_rows::Int64
    main_ms::Float64
end

# ---- enums ---------------------------------------------------------------------------------------------------------
const ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON = Int32(0), Int32(1), Int32(2), Int32(3)
const PROC_NULL, PROC_COULOMB, PROC_RBEB, PROC_MOLLER, PROC_BHABA, PROC_SELTZER, PROC_COMPTON, PROC_PHOTOELECTRIC,
      PROC_BETHE_HEITLER, PROC_ANIHILATION, PROC_LX_EXCITATION, PROC_LX_IONIZATION, PROC_LX_ATTACHMENT, PROC_LX_ELASTIC =
      Int32.(0:13)
const ERR_BITS = (1 => "CAPACITY_OVERFLOW (population.jl:107)", 2 => "RATE_BOUND_VIOLATED (collisions.jl:186)",
                  4 => "ENERGY_OUT_OF_TABLE (collision_table.jl:91)", 8 => "NAN_STATE",
                  16 => "SAMPLER_INVARIANT (rbeb.jl:63, seltzer.jl:73, photo_electric.jl:71)")

species(::Type{<:ElectronState}) = ELECTRON
species(::Type{<:PhotonState}) = PHOTON
species(::Type{<:PositronState}) = POSITRON
species(::Type{Particulator.SlowElectronState{T}}) where T = SLOW_ELECTRON
species(::Type{Electron}) = ELECTRON
species(::Type{Photon}) = PHOTON
species(::Type{Positron}) = POSITRON

# ---- context -------------------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    tables::IdDict{Any,Int32}       # host table object -> device table id
    cheb_losses::IdDict{Any,Int32}
    sb::IdDict{Any,Int32}
end

"""
    Context(device = 0)

One library context on CUDA device `device` (one context per GPU, one host task per context).  Throws unless a
compute-capability-10.x device is present: there is no CPU fallback.
"""
function Context(device::Integer = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ptl_context_create, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, out)
    rc == 0 || error("ptl_context_create: status $rc" * (rc == -2 ? " (no sm_100 device: there is no CPU fallback)" : ""))
    ctx = Context(out[], IdDict{Any,Int32}(), IdDict{Any,Int32}(), IdDict{Any,Int32}())
    finalizer(c -> (c.h == C_NULL || ccall((:ptl_context_destroy, LIB), Int32, (Ptr{Cvoid},), c.h); c.h = C_NULL), ctx)
    return ctx
end

last_error(ctx::Context) = unsafe_string(ccall((:ptl_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))

"Usage errors (negative) throw ErrorException; sticky device conditions (positive) throw the AssertionError the reference would."
function check(ctx::Context, rc::Integer, what)
    rc < 0 && error("$what: status $rc: $(last_error(ctx))")
    if rc > 0
        names = join((name for (bit, name) in ERR_BITS if rc & bit != 0), " | ")
        throw(AssertionError("$what: device condition $names"))
    end
    return rc
end

"`ptl_set_rng`: the device path does not use Julia's RNG; streams are Philox4x32-10 keyed by particle uid."
set_rng!(ctx::Context, seed::Integer, step::Integer = 0) =
    check(ctx, ccall((:ptl_set_rng, LIB), Int32, (Ptr{Cvoid}, UInt64, UInt32), ctx.h, seed, step), "set_rng!")
function get_rng(ctx::Context)
    s, st = Ref{UInt64}(0), Ref{UInt32}(0)
    ccall((:ptl_get_rng, LIB), Int32, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt32}), ctx.h, s, st)
    return (seed = s[], step = st[])
end
uid_counter(ctx::Context) = ccall((:ptl_get_uid_counter, LIB), UInt64, (Ptr{Cvoid},), ctx.h)
set_uid_counter!(ctx::Context, v::Integer) = check(ctx, ccall((:ptl_set_uid_counter, LIB), Int32, (Ptr{Cvoid}, UInt64), ctx.h, v), "set_uid_counter!")
error_flags(ctx::Context; clear = false) = ccall((:ptl_error_flags, LIB), Int32, (Ptr{Cvoid}, Int32), ctx.h, clear ? 1 : 0)
synchronize(ctx::Context) = check(ctx, ccall((:ptl_synchronize, LIB), Int32, (Ptr{Cvoid},), ctx.h), "synchronize")
set_option!(ctx::Context, name::AbstractString, value::Integer) =
    check(ctx, ccall((:ptl_set_option, LIB), Int32, (Ptr{Cvoid}, Cstring, Int64), ctx.h, name, value), "set_option!")
