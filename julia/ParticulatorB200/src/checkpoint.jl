# checkpoint.jl — restart state (SURVEY 8 f3): every column of every population (uid included: it keys the RNG streams), the
# simulation time, and the context's (seed, advance-call index, uid counter).  Stored with Julia's Serialization; the Python
# host writes the same content as an .npz (particulator.jl_b200/checkpoint.py).
using Serialization

function save_checkpoint(path::AbstractString, mp::DeviceMultiPopulation, t::Real)
    pops = Dict{Symbol,Any}()
    for (name, d) in pairs(mp)
        uid = download!(d)
        n = d.host.n[]
        pops[name] = (particles = d.host.particles[1:n], uid = uid)
    end
    rng = get_rng(mp.ctx)
    serialize(path, (format = "particulator_b200.checkpoint", version = 1, t = Float64(t), seed = rng.seed, step = rng.step,
                     next_uid = uid_counter(mp.ctx), populations = pops))
end

function load_checkpoint!(path::AbstractString, mp::DeviceMultiPopulation)
    ck = deserialize(path)
    ck.format == "particulator_b200.checkpoint" || error("not a particulator_b200 checkpoint")
    for (name, d) in pairs(mp)
        haskey(ck.populations, name) || error("the checkpoint has no population $name")
        saved = ck.populations[name]
        n = length(saved.particles)
        n <= length(d.host.particles) || error("population $name: $n saved rows exceed the capacity")
        d.host.particles[1:n] = saved.particles
        d.host.n[] = n
        upload!(d; uid = Vector{UInt64}(saved.uid))
    end
    set_rng!(mp.ctx, ck.seed, ck.step)
    set_uid_counter!(mp.ctx, max(ck.next_uid, uid_counter(mp.ctx)))
    return ck.t
end
