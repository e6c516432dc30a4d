"""
    ParticulatorB200

Device twin of the particle-advance hot path of Particulator.jl.  The host keeps the reference's public surface
(`src/Particulator.jl:2-18`): particle definitions, collision-process types, table builders, pushers, fields,
callbacks and `run!`.  `DevicePopulation` / `DeviceMultiPopulation` replace `Population` / `MultiPopulation`
(`src/population.jl:7-44`, `src/mixed_population.jl:4-18`), and the generic functions `run!` and the callbacks call on
them (`advance!`, `init!`, `droplow!`, `repack!`, `nparticles`, `nactives`, `weight`, `meanenergy`, `maxenergy`,
`spread`, `posvar`, `roulette!`, `split!`, `shuffle!`, `add_particle!`, `remove_particle!`, `empty!`) become `ccall`s
into `libparticulator_b200.so` (C ABI: `include/particulator_b200.h`).  No CUDA.jl, no CPU fallback: creating a
`Context` fails unless a compute-capability-10.x device is present.

    using Particulator, ParticulatorB200
    ctx   = ParticulatorB200.Context(0)
    dmp   = DeviceMultiPopulation(ctx, mpopl)            # uploads tables and particles
    run!(dmp, RK2Pusher(ElectromagneticField(efield, bfield)), tfinal, dt, callback)
    download!(dmp)                                       # back into the host StructArrays

STATUS: written against the header and the reference sources; Julia is not installed in the image this repository is
built in, so this package has NOT been executed there.  The same ABI is exercised end to end from Python
(`particulator.jl_b200/_lib.py`) by the GPU test-suite.
"""
module ParticulatorB200

using StaticArrays, StructArrays, Random
import Particulator
import Particulator: Population, MultiPopulation, ParticleState, ParticleType, ElectronState, PhotonState, PositronState,
                     Electron, Photon, Positron,
                     ChebyshevCollisionTable, CollisionTable, LogLinRange, NullCollision,
                     RK2Pusher, NullPusher, RestrictedPusher, NullForcing, CombinedForcing, RestrictedForcing,
                     ElectromagneticField, HomogeneousField, DoubleLayerField, StepField, ConfinedDoubleLayerField,
                     ContinuumLoss, ChebContinuumLoss,
                     AbstractCallback, VoidCallback, CombinedCallback, WallCallback, CollisionCounter,
                     RelativisticCoulomb, RBEB, Moller, Bhaba, SeltzerBerger, Compton, KleinNishinaCompton,
                     PhotoElectric, BetheHeitler, PositronAnihilation,
                     advance!, init!, droplow!, repack!, nparticles, nactives, weight, meanenergy, maxenergy, spread,
                     posvar, roulette!, split!, shuffle!, add_particle!, remove_particle!, kinenergy

export Context, DevicePopulation, DeviceMultiPopulation, upload!, download!, set_rng!, get_rng, error_flags,
       wall_records, collision_counts, histogram, advance_stats,
       comm_unique_id, comm_init!, comm_destroy!, diag_allreduce, histogram_allreduce, rebalance!,
       save_checkpoint, load_checkpoint!

include("capi.jl")
include("tables.jl")
include("populations.jl")
include("descriptors.jl")
include("hotpath.jl")
include("comm.jl")
include("checkpoint.jl")

end # module
