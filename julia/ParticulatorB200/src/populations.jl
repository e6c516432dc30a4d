# populations.jl — DevicePopulation / DeviceMultiPopulation: the device twins of Population (population.jl:7-44) and
# MultiPopulation (mixed_population.jl:4-18).  The host Population stays attached: upload!/download! move the StructArray
# columns as they lie in memory (x and p are Vector{SVector{3,Float64}}: xyz-interleaved, exactly what the ABI takes).

struct DevicePopulation{PS}
    ctx::Context
    id::Int32
    table::Int32
    host::Population
end

const _PF = Ptr{Float64}

function DevicePopulation(ctx::Context, popl::Population{PS}; capacity::Integer = length(popl.particles)) where PS
    tab = device_table(ctx, popl.collisions)
    id = ccall((:ptl_population_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Float64, Int32),
               ctx.h, species(PS), capacity, Float64(popl.energy_cut), tab)
    id < 0 && error("ptl_population_create: status $id: $(last_error(ctx))")
    d = DevicePopulation{PS}(ctx, id, tab, popl)
    upload!(d)
    return d
end

# positions / velocities of a SlowElectronState live in fields x / v (slow-electron.jl:9-17); the device column is "p"
_pcol(p, ::Type{PS}) where PS = hasproperty(p, :p) ? p.p : p.v

"Host StructArray -> device columns (sets n).  uids are assigned by the context unless `uid` is given."
function upload!(d::DevicePopulation{PS}; uid::Union{Nothing,Vector{UInt64}} = nothing) where PS
    p = d.host.particles
    n = d.host.n[]
    pc = _pcol(p, PS)
    t = hasproperty(p, :t) ? p.t : zeros(Float64, length(p))          # the stale slow-electron state has no t / r (Appendix B)
    r = hasproperty(p, :r) ? p.r : zeros(Float64, length(p))
    act = Vector{UInt8}(p.active[1:n])
    GC.@preserve p pc t r act uid begin
        rc = ccall((:ptl_population_upload, LIB), Int32,
                   (Ptr{Cvoid}, Int32, Int64, _PF, _PF, _PF, _PF, _PF, _PF, Ptr{UInt8}, Ptr{UInt64}),
                   d.ctx.h, d.id, n, pointer(p.x), pointer(pc), pointer(p.w), pointer(t), pointer(p.s), pointer(r),
                   pointer(act), uid === nothing ? C_NULL : pointer(uid))
        check(d.ctx, rc, "upload!")
    end
    return d
end

"Device columns -> host StructArray (sets popl.n).  Returns the uid column."
function download!(d::DevicePopulation{PS}) where PS
    p = d.host.particles
    cap = length(p)
    pc = _pcol(p, PS)
    t = hasproperty(p, :t) ? p.t : Vector{Float64}(undef, cap)
    r = hasproperty(p, :r) ? p.r : Vector{Float64}(undef, cap)
    act = Vector{UInt8}(undef, cap)
    uid = Vector{UInt64}(undef, cap)
    n = GC.@preserve p pc t r act uid ccall((:ptl_population_download, LIB), Int64,
            (Ptr{Cvoid}, Int32, Int64, _PF, _PF, _PF, _PF, _PF, _PF, Ptr{UInt8}, Ptr{UInt64}),
            d.ctx.h, d.id, cap, pointer(p.x), pointer(pc), pointer(p.w), pointer(t), pointer(p.s), pointer(r),
            pointer(act), pointer(uid))
    n < 0 && error("ptl_population_download: status $n: $(last_error(d.ctx))")
    for i in 1:n
        p.active[i] = act[i] != 0
    end
    d.host.n[] = n
    return resize!(uid, n)
end

struct DeviceMultiPopulation{NT<:NamedTuple}
    ctx::Context
    id::Int32
    index::NT                     # same names, same order as MultiPopulation.index: the processing order of advance1!
end

function DeviceMultiPopulation(ctx::Context, mp::MultiPopulation)
    names = keys(mp.index)
    pops = map(p -> DevicePopulation(ctx, p), Tuple(mp.index))
    ids = Int32[p.id for p in pops]
    id = GC.@preserve ids ccall((:ptl_multipop_create, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32), ctx.h, ids, length(ids))
    id < 0 && error("ptl_multipop_create: status $id: $(last_error(ctx))")
    return DeviceMultiPopulation(ctx, id, NamedTuple{names}(pops))
end

Base.get(mp::DeviceMultiPopulation, ::Type{ParticleType{S}}) where S = getfield(mp.index, S)     # mixed_population.jl:15
Base.map(f, mp::DeviceMultiPopulation) = map(f, Tuple(mp.index))
Base.foreach(f, mp::DeviceMultiPopulation) = foreach(f, Tuple(mp.index))
Base.pairs(mp::DeviceMultiPopulation) = pairs(mp.index)
download!(mp::DeviceMultiPopulation) = map(download!, mp)
upload!(mp::DeviceMultiPopulation) = (foreach(upload!, mp); mp)

# ---- generic functions of population.jl on the device twin ---------------------------------------------------------------
function _diag(d::DevicePopulation)
    o = Ref{DiagOut}()
    check(d.ctx, ccall((:ptl_diag, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{DiagOut}), d.ctx.h, d.id, o), "diag")
    return o[]
end
Particulator.nparticles(d::DevicePopulation) = Int(ccall((:ptl_population_n, LIB), Int64, (Ptr{Cvoid}, Int32), d.ctx.h, d.id))   # population.jl:78
Particulator.nactives(d::DevicePopulation) = Int(_diag(d).nactive)                                                            # :89-97
Particulator.weight(d::DevicePopulation) = _diag(d).weight                                                                    # :130-140
Particulator.meanenergy(d::DevicePopulation) = (o = _diag(d); o.wenergy / o.weight)                                           # :152-166
Particulator.maxenergy(d::DevicePopulation) = _diag(d).maxenergy                                                              # :172-174
function Particulator.spread(d::DevicePopulation)                                                                             # :180-203
    o = _diag(d)
    xm = SVector(o.wx) / o.weight
    return xm, sqrt(abs(o.wr2 / o.weight - xm' * xm))
end
Particulator.posvar(d::DevicePopulation) = (o = _diag(d); SVector(o.wx2) ./ o.weight .- (SVector(o.wx) ./ o.weight) .^ 2)      # :205-223
Base.length(d::DevicePopulation) = nparticles(d)
Base.empty!(d::DevicePopulation) = (check(d.ctx, ccall((:ptl_population_clear, LIB), Int32, (Ptr{Cvoid}, Int32), d.ctx.h, d.id), "empty!"); d)   # :69

"droplow!(popl, thres) population.jl:273-284: flag E < thres, then the same tail-fill compaction as repack!."
function Particulator.droplow!(d::DevicePopulation, thres = 0.0)
    n = ccall((:ptl_droplow, LIB), Int64, (Ptr{Cvoid}, Int32, Float64), d.ctx.h, d.id, Float64(thres))
    n < 0 && error("droplow!: status $n: $(last_error(d.ctx))")
    return nothing
end
function Particulator.repack!(d::DevicePopulation)                                                                            # :229-259
    n = ccall((:ptl_repack, LIB), Int64, (Ptr{Cvoid}, Int32), d.ctx.h, d.id)
    n < 0 && error("repack!: status $n: $(last_error(d.ctx))")
    return nothing
end

# roulette! / split!: population.jl:291-340.  Numbers go straight through; a function of the energy is tabulated on `nodes`
# log-spaced nodes between the energy cut and `emax` and interpolated linearly by the library (a closure cannot cross the ABI).
function _law(f, d::DevicePopulation, emax, nodes)
    lo = log10(max(Float64(d.host.energy_cut), 1.602e-22))
    hi = log10(Float64(emax))
    x = range(lo, hi; length = nodes)
    return lo, hi, Float64[f(10.0^q) for q in x]
end
Particulator.roulette!(p::Number, d::DevicePopulation) =
    check(d.ctx, ccall((:ptl_roulette, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), d.ctx.h, d.id, Float64(p)), "roulette!")
function Particulator.roulette!(f::Function, d::DevicePopulation; emax = 1.602e-10, nodes = 1024)
    lo, hi, v = _law(f, d, emax, nodes)
    check(d.ctx, GC.@preserve(v, ccall((:ptl_roulette_law, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, Float64, Int32, Int32, _PF),
                                       d.ctx.h, d.id, lo, hi, length(v), 1, v)), "roulette!")
end
Particulator.split!(p::Number, d::DevicePopulation) =
    check(d.ctx, ccall((:ptl_split, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), d.ctx.h, d.id, Float64(p)), "split!")
function Particulator.split!(f::Function, d::DevicePopulation; emax = 1.602e-10, nodes = 1024)
    lo, hi, v = _law(f, d, emax, nodes)
    check(d.ctx, GC.@preserve(v, ccall((:ptl_split_law, LIB), Int32, (Ptr{Cvoid}, Int32, Float64, Float64, Int32, Int32, _PF),
                                       d.ctx.h, d.id, lo, hi, length(v), 1, v)), "split!")
end
Particulator.shuffle!(d::DevicePopulation) = check(d.ctx, ccall((:ptl_shuffle, LIB), Int32, (Ptr{Cvoid}, Int32), d.ctx.h, d.id), "shuffle!")   # :266-271

"add_particle!(popl, state) population.jl:103-113 (slow path: one particle, host-synchronous)."
function Particulator.add_particle!(d::DevicePopulation{PS}, s::PS) where PS
    x = Float64[s.x...]
    p = Float64[(hasproperty(s, :p) ? s.p : s.v)...]
    t = hasproperty(s, :t) ? s.t : 0.0
    r = hasproperty(s, :r) ? s.r : 0.0
    j = GC.@preserve x p ccall((:ptl_population_append, LIB), Int64, (Ptr{Cvoid}, Int32, _PF, _PF, Float64, Float64, Float64, Float64, UInt64),
                               d.ctx.h, d.id, x, p, s.w, t, s.s, r, 0)
    j == -7 && throw(AssertionError("add_particle!: population is full (population.jl:107)"))
    j < -1 && error("add_particle!: status $j: $(last_error(d.ctx))")
    return j + 1                                     # 1-based row, 0 = below the energy cut
end
Particulator.remove_particle!(d::DevicePopulation, i::Integer) =                                                               # :120-122
    check(d.ctx, ccall((:ptl_population_deactivate, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), d.ctx.h, d.id, i - 1), "remove_particle!")

"Weighted histogram of the kinetic energy (:energy) or of cos(theta_z) (:costheta) over the active particles (scripts/beam.jl:137-146)."
function histogram(d::DevicePopulation, quantity::Symbol, lo, hi, nbins; logscale = false)
    out = zeros(Float64, nbins)
    q = quantity === :energy ? 0 : quantity === :costheta ? 1 : error("quantity must be :energy or :costheta")
    check(d.ctx, GC.@preserve(out, ccall((:ptl_histogram, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Float64, Int32, Int32, _PF),
                                         d.ctx.h, d.id, q, Float64(lo), Float64(hi), nbins, logscale ? 1 : 0, out)), "histogram")
    return out
end
