"""particulator_b200 — B200-native particle-advance hot path behind the API of aluque/Particulator.jl.

Host-side mirror of the reference's public surface for this path (src/Particulator.jl:2-18 exports):
particle species, collision-process types, table builders, Population / MultiPopulation, pushers,
fields, callbacks, `advance`, `droplow`, `repack`, `run` and the diagnostics.  All per-particle work
happens in hand-written sm_100a CUDA kernels (csrc/) reached through the C ABI of
include/particulator_b200.h; there is no CPU fallback in this package."""
from . import constants as co
from .processes import (ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON, NullCollision, RelativisticCoulomb, RBEB, Moller,
                        Bhaba, Compton, KleinNishinaCompton, PhotoElectric, BetheHeitler, PositronAnihilation,
                        SeltzerBerger, Excitation, Ionization, Attachment, Elastic, ORBITALS, N2_ORBITALS, O2_ORBITALS,
                        speed)
from .cheby import BinaryIntervals, chebfit, chebeval, chebdiff, chebval, precheb
from .tables import (ChebyshevCollisionTable, CollisionTable, collision_table_from_processes, compratebound,
                     air_composition, build_electron_collision_table, build_positron_collision_table,
                     build_photon_collision_table, synthetic_lxcat_table, lxcat_table_from_rates, loglinrange)
from . import seltzer
from .lxcat import load_lxcat, lxcat_collision_table, ensure_elastic
from ._lib import PtlError, Backend, cuda_backend, LIB_PATH, ABI_SYMBOLS, ABI_SYMBOLS_CORE
from .context import Context
from .field import (ZeroField, HomogeneousField, DoubleLayerField, StepField, ConfinedDoubleLayerField,
                    ElectromagneticField)
from .pusher import (NullForcing, CombinedForcing, RestrictedForcing, RK2Pusher, NullPusher, RestrictedPusher,
                     ContinuumLoss, ChebContinuumLoss)
from .population import (Population, kinenergy, momentum_norm_from_kin, nparticles, nactives, weight, meanenergy,
                         maxenergy, spread, posvar, empty, add_particle, remove_particle, repack, droplow, roulette,
                         split, shuffle)
from .mixed_population import MultiPopulation, init, advance, last_advance_stats
from .callback import (AbstractCallback, VoidCallback, CombinedCallback, CollisionCounter, WallCallback,
                       ParticleCountCallback, RouletteCallback, SplitCallback, PopulationTargetCallback)
from .run import run
from . import checkpoint
from .checkpoint import save_checkpoint, load_checkpoint, read_checkpoint

Electron, Photon, Positron, SlowElectron = ELECTRON, PHOTON, POSITRON, SLOW_ELECTRON
