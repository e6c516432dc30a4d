# Smoke test of the device twin against the host reference on a B200: the same small beam (scripts/beam.jl) advanced by
# Particulator.jl on the CPU and by ParticulatorB200 on the GPU must agree statistically (the RNGs differ: Xoshiro vs
# uid-keyed Philox), and the deterministic pieces (table lookups, droplow!/repack! permutation) must agree exactly.
using Test, Particulator, ParticulatorB200, StaticArrays, Random
import Particulator: co

comp = Dict("N2" => co.nair * 0.79, "O2" => co.nair * 0.21)
dt, efield = 2.5e-11, 5e5
Fdt = co.elementary_charge * efield * dt
eproc = [(2 * comp["N2"], RelativisticCoulomb(7)), (2 * comp["O2"], RelativisticCoulomb(8)),
         (2 * comp["N2"], SeltzerBerger(7)), (2 * comp["O2"], SeltzerBerger(8)),
         [(comp["N2"], orb) for orb in ORBITALS["N2"]]..., [(comp["O2"], orb) for orb in ORBITALS["O2"]]...]
ecolls = collision_table_from_processes(eproc, Electron, Fdt; safety = 1.15)
gcolls = collision_table_from_processes([(2 * comp["N2"], PhotoElectric(7)), (2 * comp["O2"], PhotoElectric(8)),
                                         (2 * comp["N2"], BetheHeitler(7)), (2 * comp["O2"], BetheHeitler(8)),
                                         (2 * comp["N2"], Compton(7)), (2 * comp["O2"], Compton(8))], Photon, 0; safety = 1.15)
pcolls = collision_table_from_processes([(2 * comp["N2"], RelativisticCoulomb(7)), (2 * comp["O2"], RelativisticCoulomb(8)),
                                         (2 * comp["N2"], Bhaba(7, 1e2 * co.eV)), (2 * comp["O2"], Bhaba(8, 1e2 * co.eV)),
                                         (2 * comp["N2"], PositronAnihilation(7)), (2 * comp["O2"], PositronAnihilation(8))],
                                        Positron, Fdt; safety = 1.15)

function world(n)
    pn = Particulator.momentum_norm_from_kin(Electron, 3e6 * co.eV)
    init = [ElectronState(SA[0.0, 0.0, 0.0], SA[0.0, 1e-6 * pn, pn]) for _ in 1:n]
    MultiPopulation(:electron => Population(40n, init, ecolls, 1e3 * co.eV),
                    :photon => Population(40n, PhotonState{Float64}[], gcolls, 1e3 * co.eV),
                    :positron => Population(4n, PositronState{Float64}[], pcolls, 1e2 * co.eV))
end
pusher = RK2Pusher(ElectromagneticField(HomogeneousField(SA[0.0, 0.0, -efield]), HomogeneousField(SA[0.0, 0.0, 0.0])))

@testset "ParticulatorB200" begin
    ctx = ParticulatorB200.Context(0)
    @testset "table lookups are bit-exact" begin
        e = exp.(range(log(1e-2 * co.eV), log(0.99 * ecolls.b.xmax); length = 500))
        rates, bound = ParticulatorB200.table_eval(ctx, ecolls, e)
        for (i, x) in enumerate(e)
            pre = Particulator.presample(ecolls, nothing, x)
            @test all(rates[j, i] === Particulator.rate(ecolls, j, pre) for j in 1:length(ecolls.proc))
            @test bound[i] === Particulator.ratebound(ecolls, x)
        end
    end
    @testset "run! on the device twin" begin
        Random.seed!(1)
        host = world(2000); init!(host)
        run!(host, pusher, 40dt, dt, VoidCallback(); output_dt = nothing, verbosity = 0)
        dev = DeviceMultiPopulation(ctx, world(2000)); set_rng!(ctx, 1, 0); init!(dev)
        run!(dev, pusher, 40dt, dt, VoidCallback(); output_dt = nothing, verbosity = 0)
        eh, ed = get(host, Electron), get(dev, Electron)
        @test isapprox(nparticles(ed), nparticles(eh); rtol = 0.1)
        @test isapprox(Particulator.meanenergy(ed), Particulator.meanenergy(eh); rtol = 0.05)
        @test isapprox(Particulator.spread(ed)[1][3], Particulator.spread(eh)[1][3]; rtol = 0.05)
    end
end
