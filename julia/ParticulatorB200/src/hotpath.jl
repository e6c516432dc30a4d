# hotpath.jl — advance! and what run! does around it (mixed_population.jl:38-47, run.jl:1-29).

"init!(mpopl) mixed_population.jl:20-35: setr! on all actives."
Particulator.init!(mp::DeviceMultiPopulation) =
    check(mp.ctx, ccall((:ptl_init, LIB), Int32, (Ptr{Cvoid}, Int32), mp.ctx.h, mp.id), "init!")

"""
    advance!(mpopl::DeviceMultiPopulation, pusher, tfinal, callback = VoidCallback())

The hot call (mixed_population.jl:38-47): every active particle of every population is carried to `tfinal` on the device
(advance_init!, the sub-step loop of advance1!, the pusher, do_one_collision!, collide, apply!, add_particle!).  Wall
crossings recorded by `WallCallback`s are appended to their `accum` vectors and `CollisionCounter` totals are updated
after the call, so scripts that read them keep working.
"""
function Particulator.advance!(mp::DeviceMultiPopulation, pusher, tfinal, callback = VoidCallback())
    pd = Ref(pusherdesc(mp.ctx, pusher))
    cbd, walls, counter = callbackdesc(callback)
    has_cb = cbd.nwalls > 0 || cbd.count_collisions != 0
    cb = Ref(cbd)
    rc = GC.@preserve pd cb ccall((:ptl_advance, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{PusherDesc}, Float64, Ptr{CallbackDesc}),
                                  mp.ctx.h, mp.id, pd, Float64(tfinal), has_cb ? Base.unsafe_convert(Ptr{CallbackDesc}, cb) : C_NULL)
    check(mp.ctx, rc, "advance!")
    for (k, w) in enumerate(walls)
        append!(w.accum, wall_records(mp, k, w; clear = true))
    end
    counter === nothing || _fold_counts!(counter, mp)
    return nothing
end

"Statistics of the last advance! call (passes of advance1!, sub-steps, rows visited, births, kernels launched)."
function advance_stats(mp::DeviceMultiPopulation)
    o = Ref{AdvanceStats}()
    ccall((:ptl_last_advance_stats, LIB), Int32, (Ptr{Cvoid}, Ref{AdvanceStats}), mp.ctx.h, o)
    return o[]
end

# WallCallback.accum read-out (callback.jl:146-184): states interpolated at the wall, as lincomb builds them
function wall_records(mp::DeviceMultiPopulation, iwall::Integer, w::WallCallback{P}; clear = true) where P
    n = ccall((:ptl_wall_records, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, _PF, _PF, _PF, _PF, Int32),
              mp.ctx.h, iwall - 1, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0)
    n <= 0 && return P[]
    x, p = Matrix{Float64}(undef, 3, n), Matrix{Float64}(undef, 3, n)
    wgt, t = Vector{Float64}(undef, n), Vector{Float64}(undef, n)
    GC.@preserve x p wgt t ccall((:ptl_wall_records, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, _PF, _PF, _PF, _PF, Int32),
                                 mp.ctx.h, iwall - 1, n, x, p, wgt, t, clear ? 1 : 0)
    return P[P(SVector{3}(x[:, i]), SVector{3}(p[:, i]), wgt[i], t[i]) for i in 1:n]
end

# CollisionCounter (callback.jl:118-141): per-process totals keyed by the process TYPE, like the reference's Dict
function collision_counts(d::DevicePopulation; clear = false)
    np = length(d.host.collisions.proc)
    c = zeros(Int64, np + 1)
    check(d.ctx, GC.@preserve(c, ccall((:ptl_collision_counts, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int32), d.ctx.h, d.table, c, clear ? 1 : 0)),
          "collision_counts")
    return c
end
function _fold_counts!(cc::CollisionCounter, mp::DeviceMultiPopulation)
    foreach(mp) do d
        c = collision_counts(d; clear = true)
        procs = d.host.collisions.proc
        for j in eachindex(c)
            c[j] == 0 && continue
            key = j <= length(procs) ? typeof(procs[j]) : NullCollision
            if haskey(cc.d, key)
                cc.d[key][] += c[j]
            else
                cc.d[key] = Base.Threads.Atomic{Int}(c[j])
            end
        end
    end
end

# run!(mpopl, pusher, tfinal, dt, callback) (run.jl:1-29) works unchanged on a DeviceMultiPopulation: it only calls
# advance!, droplow!, onstep, onoutput, nparticles and spread, all defined above.  The between-step callbacks
# (RouletteCallback, SplitCallback, PopulationTargetCallback, ParticleCountCallback: callback.jl:188-268) call
# nactives / weight / roulette! / split! / repack! on `pairs(mpopl)` / `get(mpopl, ParticleType{S})`, also defined above.
