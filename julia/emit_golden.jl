#!/usr/bin/env julia
# emit_golden.jl — run the REFERENCE (aluque/Particulator.jl, unmodified) on fixed inputs and write what it computes, in the
# schema tests/golden/import_reference_vectors.py packs into tests/golden/reference_vectors.npz.  tests/test_reference_vectors.py
# activates when that file exists and compares the CPU oracle (and, on a GPU box, the CUDA path) against it: this is the step
# that clears the "parity unpinned" flag of oracle/ptl_oracle.c, and it needs a Julia runtime, which the build image lacks.
#
#   julia --project=/path/to/Particulator.jl julia/emit_golden.jl OUTDIR
#
# Everything emitted is deterministic:
#   tables_*      Chebyshev coefficient arrays of the three air tables of scripts/beam.jl:94-129 (pins chebfit, every
#                 totalcs, compratebound, the process ordering) and rate / ratebound lookups at fixed energies
#   sb_*          the Seltzer-Berger inverse-CDF tables and totalcs of N and O (seltzer.jl:23-49,163-231)
#   kin_*, push_* kinematics and one RK2 push (electron.jl:47-56, pusher.jl:41-63) of fixed states
#   turn_*        util.jl:40-57
#   collide_*     every `collide` of the three tables on fixed momenta with the uniforms it consumed: the global RNG is seeded,
#                 64 scalar rand() values are recorded, the RNG is re-seeded and collide() runs — so the recorded values ARE its
#                 draws, in order.  The oracle replays them with ora_collide_replay.
#   repack_*      the permutation of repack! (population.jl:229-259) for fixed active patterns
using Particulator, StaticArrays, Random, Printf
import Particulator: co, collide, kinenergy, velocity, momentum_norm_from_kin, advance_particle, presample, rate, ratebound,
                     speed, totalcs, turn
import JSON

outdir = length(ARGS) >= 1 ? ARGS[1] : "reference_vectors"
mkpath(outdir)
manifest = Dict{String,Any}()
function emit(name, a::AbstractArray{T}) where T <: Union{Float64,Int64,UInt8}
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, Array(a))                                   # column-major; the importer transposes by the recorded shape
    end
    manifest[name] = Dict("shape" => collect(size(a)), "dtype" => string(T), "order" => "F")
end

# ---- tables (scripts/beam.jl:94-129) -------------------------------------------------------------------------------------
const DT, EFIELD, SAFETY = 2.5e-11, 5e5, 1.15
comp = Dict("N2" => co.nair * 0.79, "O2" => co.nair * 0.21)
Fdt = co.elementary_charge * EFIELD * DT
eproc = [(2 * comp["N2"], RelativisticCoulomb(7)), (2 * comp["O2"], RelativisticCoulomb(8)),
         (2 * comp["N2"], SeltzerBerger(7)), (2 * comp["O2"], SeltzerBerger(8)),
         [(comp["N2"], orb) for orb in ORBITALS["N2"]]..., [(comp["O2"], orb) for orb in ORBITALS["O2"]]...]
pproc = [(2 * comp["N2"], RelativisticCoulomb(7)), (2 * comp["O2"], RelativisticCoulomb(8)),
         (2 * comp["N2"], Bhaba(7, 1e2 * co.eV)), (2 * comp["O2"], Bhaba(8, 1e2 * co.eV)),
         (2 * comp["N2"], PositronAnihilation(7)), (2 * comp["O2"], PositronAnihilation(8))]
gproc = [(2 * comp["N2"], PhotoElectric(7)), (2 * comp["O2"], PhotoElectric(8)), (2 * comp["N2"], BetheHeitler(7)),
         (2 * comp["O2"], BetheHeitler(8)), (2 * comp["N2"], Compton(7)), (2 * comp["O2"], Compton(8))]
tables = Dict("electron" => (collision_table_from_processes(eproc, Electron, Fdt; safety = SAFETY), Electron, ElectronState),
              "positron" => (collision_table_from_processes(pproc, Positron, Fdt; safety = SAFETY), Positron, PositronState),
              "photon" => (collision_table_from_processes(gproc, Photon, 0; safety = SAFETY), Photon, PhotonState))

golden_energies(xmax) = vcat(exp.(range(log(1e-2 * co.eV), log(0.99 * xmax); length = 97)), [1e3 * co.eV, 0.0])
function golden_momenta(P, lo_eV; n = 16)
    K = exp.(range(log(lo_eV), log(1.5e8); length = n)) .* co.eV
    pn = momentum_norm_from_kin.(Ref(P), K)
    ang = range(0.3, 2.7; length = n)
    return [SA[sin(a) * cos(2.1a), sin(a) * sin(2.1a), cos(a)] * q for (a, q) in zip(ang, pn)]
end
const LOW = Dict("electron" => 1.2e3, "positron" => 3e2, "photon" => 1.2e3)

outcome_row(o::NullOutcome) = (0, Any[])
outcome_row(o::StateChangeOutcome) = (1, Any[o.state])
outcome_row(o::NewParticleOutcome) = (2, Any[o.state1, o.state2])
outcome_row(o::RemoveParticleOutcome) = (3, Any[])
outcome_row(o::ReplaceParticleOutcome) = (4, Any[nothing, o.state2])
outcome_row(o::ReplaceParticlePairOutcome) = (5, Any[nothing, o.state2, o.state3])
spid(s::ElectronState) = 0; spid(s::PhotonState) = 1; spid(s::PositronState) = 2

for (name, (tab, P, PS)) in tables
    emit("tables_$(name)_rate", tab.rate)                                  # [order, nprocs, k+1]
    emit("tables_$(name)_ratebound", tab.ratebound)                        # [order, k+1]
    emit("tables_$(name)_prockind", Int64[findfirst(==(nameof(typeof(p))), (:NullCollision, :RelativisticCoulomb, :RBEB, :Moller, :Bhaba,
          :SeltzerBerger, :Compton, :PhotoElectric, :BetheHeitler, :PositronAnihilation)) - 1 for p in tab.proc])
    emit("tables_$(name)_procZ", Float64[hasproperty(p, :Z) ? p.Z : (hasproperty(p, :B) ? p.B : -1.0) for p in tab.proc])
    e = golden_energies(tab.b.xmax)
    emit("lookup_$(name)_energy", e)
    emit("lookup_$(name)_rates", Float64[rate(tab, j, presample(tab, nothing, x)) for j in 1:length(tab.proc), x in e])
    emit("lookup_$(name)_bound", Float64[ratebound(tab, x) for x in e])
    for (j, proc) in enumerate(tab.proc)
        lo = LOW[name]
        proc isa BetheHeitler && (lo = 1.05e6)
        proc isa RBEB && (lo = max(lo, 1.05 * proc.B / co.eV))
        moms = golden_momenta(P, lo)
        U = zeros(64, length(moms))
        out = zeros(24, length(moms))
        for (i, p) in enumerate(moms)
            seed = 1000 * j + i
            Random.seed!(seed)
            U[:, i] = [rand() for _ in 1:64]
            Random.seed!(seed)
            st = PS(SA[0.0, 0.0, 0.0], p, 1.0, 0.0, 1.0, 0.0)          # explicit s: the constructor draws nothing
            kind, states = outcome_row(collide(proc, st, kinenergy(st)))
            out[1, i] = kind
            for (q, s) in enumerate(states)
                s === nothing && continue
                q >= 2 && (out[q, i] = spid(s))
                out[4q + 1:4q + 3, i] = s.p
                out[4q + 4, i] = s.s
            end
        end
        emit("collide_$(name)_$(j - 1)_p", reduce(hcat, moms))
        emit("collide_$(name)_$(j - 1)_uniforms", U)
        emit("collide_$(name)_$(j - 1)_out", out)
    end
end

# ---- Seltzer-Berger tables ---------------------------------------------------------------------------------------------------
for Z in (7, 8)
    sb = SeltzerBerger(Z)
    emit("sb_$(Z)_data", Matrix{Float64}(sb.data))
    emit("sb_$(Z)_log_energy", Vector{Float64}(sb.log_energy))
    K = exp.(range(log(1.1e3), log(9e9); length = 60)) .* co.eV
    emit("sb_$(Z)_K", K)
    emit("sb_$(Z)_totalcs", Float64[totalcs(sb, k) for k in K])
end

# ---- kinematics, one RK2 push, turn ------------------------------------------------------------------------------------------
moms = golden_momenta(Electron, 1.2e3; n = 32)
emit("kin_p", reduce(hcat, moms))
emit("kin_energy", Float64[kinenergy(ElectronState(SA[0.0, 0.0, 0.0], p, 1.0, 0.0, 1.0, 0.0)) for p in moms])
emit("kin_velocity", reduce(hcat, [velocity(ElectronState(SA[0.0, 0.0, 0.0], p, 1.0, 0.0, 1.0, 0.0)) for p in moms]))
psh = RK2Pusher(ElectromagneticField(HomogeneousField(SA[0.0, 0.0, -EFIELD]), HomogeneousField(SA[0.0, 0.0, 0.0])))
pushed = [advance_particle(psh, ElectronState(SA[0.1, -0.2, 0.3], p, 1.0, 0.0, 1.0, 0.0), DT) for p in moms]
emit("push_x", reduce(hcat, [s.x for s in pushed]))
emit("push_p", reduce(hcat, [s.p for s in pushed]))
cth = range(-0.95, 0.95; length = 32)
phi = range(0.1, 6.1; length = 32)
emit("turn_in", reduce(hcat, [[c, f] for (c, f) in zip(cth, phi)]))
emit("turn_out", reduce(hcat, [turn(p, c, f) for (p, c, f) in zip(moms, cth, phi)]))

# ---- repack! -----------------------------------------------------------------------------------------------------------------
for (case, (n, frac)) in enumerate([(1, 0.0), (31, 0.5), (1024, 0.3), (1025, 0.9), (5000, 0.0), (5000, 1.0), (100003, 0.1)])
    rng = MersenneTwister(n)                      # only the PATTERN matters; it is emitted alongside the result
    active = rand(rng, n) .>= frac
    states = [ElectronState(SA[Float64(i), 0.0, 0.0], SA[0.0, 0.0, 1e-21], 1.0, 0.0, 1.0, 0.0, active[i]) for i in 1:n]
    popl = Population(n, states, tables["electron"][1], 1e3 * co.eV)
    repack!(popl)
    emit("repack_$(case)_active", UInt8.(active))
    emit("repack_$(case)_order", Int64[Int(popl.particles.x[i][1]) for i in 1:nparticles(popl)])      # original (1-based) row now at row i
end

open(joinpath(outdir, "manifest.json"), "w") do io
    JSON.print(io, Dict("format" => "particulator_b200.reference_vectors", "version" => 1,
                        "julia" => string(VERSION), "arrays" => manifest), 1)
end
@printf("wrote %d arrays to %s\n", length(manifest), outdir)
