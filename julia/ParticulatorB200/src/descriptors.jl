# descriptors.jl — pushers, forcings, fields and the stock in-loop callbacks as plain C descriptors
# (pusher.jl:8-76, field.jl:4-70, continuum.jl:6-57, callback.jl:118-184).

const NOFIELD = FieldDesc(0, 0, ntuple(_ -> 0.0, 7))
const NOFORCE = ForcingDesc(0, 0, NOFIELD, NOFIELD, 0.0, 0.0, 0.0, -1, 0)

fielddesc(f::HomogeneousField) = FieldDesc(1, 0, (Float64(f.v[1]), Float64(f.v[2]), Float64(f.v[3]), 0.0, 0.0, 0.0, 0.0))                    # field.jl:4-8
fielddesc(f::DoubleLayerField) = FieldDesc(2, 0, (Float64(f.z1), Float64(f.z2), Float64(f.v[1]), Float64(f.v[2]), Float64(f.v[3]), 0.0, 0.0))   # :10-16
fielddesc(f::StepField) = FieldDesc(3, 0, (Float64(f.z), Float64.(Tuple(f.v1))..., Float64.(Tuple(f.v2))...))                                 # :22-28
fielddesc(f::ConfinedDoubleLayerField) = FieldDesc(4, 0, (Float64(f.sx), Float64(f.sy), Float64(f.sz), Float64(f.ez0), 0.0, 0.0, 0.0))           # :35-52
fielddesc(::Nothing) = NOFIELD
fielddesc(f) = error("field $(typeof(f)) is a closure or an unknown type: only the analytic fields of field.jl run on the device")

_mask(::Type{T}) where T = UInt32(1) << species(T)

forcingdesc(ctx, ::NullForcing, mask = UInt32(0)) = NOFORCE
forcingdesc(ctx, f::ElectromagneticField, mask = UInt32(0)) = ForcingDesc(1, mask, fielddesc(f.e), fielddesc(f.b), 0.0, 0.0, 0.0, -1, 0)
forcingdesc(ctx, f::ContinuumLoss, mask = UInt32(0)) = ForcingDesc(2, mask, NOFIELD, NOFIELD, Float64(f.nel), Float64(f.I), Float64(f.Tcut), -1, 0)
forcingdesc(ctx, f::ChebContinuumLoss, mask = UInt32(0)) = ForcingDesc(3, mask, NOFIELD, NOFIELD, 0.0, 0.0, 0.0, device_cheb_loss(ctx, f), 0)
forcingdesc(ctx, f::RestrictedForcing{T}, mask = UInt32(0)) where T = forcingdesc(ctx, f.forcing, _mask(T))          # pusher.jl:28-34
_terms(f::CombinedForcing) = collect(f.tpl)                                                                         # pusher.jl:14-23
_terms(f) = Any[f]

function pusherdesc(ctx, p::RK2Pusher, mask = UInt32(0))                                                            # pusher.jl:37-63
    t = ForcingDesc[forcingdesc(ctx, f) for f in _terms(p.forcing)]
    length(t) <= 4 || error("at most 4 forcing terms (PTL_MAX_FORCINGS)")
    PusherDesc(1, mask, length(t), 0, ntuple(i -> i <= length(t) ? t[i] : NOFORCE, 4))
end
pusherdesc(ctx, ::NullPusher, mask = UInt32(0)) = PusherDesc(0, mask, 0, 0, ntuple(_ -> NOFORCE, 4))               # :75-76
pusherdesc(ctx, p::RestrictedPusher{T}, mask = UInt32(0)) where T = pusherdesc(ctx, p.pusher, _mask(T))            # :67-73

# ---- in-loop callbacks.  onadvance / oncollision closures cannot run inside a kernel; the stock ones are built in. -----
const NOWALL = WallDesc(0, 0, 0.0, 0, 0)
_inloop(::VoidCallback) = Any[]
_inloop(c::CombinedCallback) = vcat((_inloop(x) for x in c.tpl)...)                                                # callback.jl:41-108
_inloop(c::Union{WallCallback,CollisionCounter}) = Any[c]
_inloop(c::AbstractCallback) = Any[]            # onstep / onoutput callbacks act between steps: nothing to send to the device

walldesc(w::WallCallback{P}) where P = WallDesc(species(P), Int32(w.coord - 1), Float64(w.v), w.drop ? 1 : 0, 0)     # callback.jl:146-164

"(descriptor, walls in descriptor order, counter or nothing) for a callback tree."
function callbackdesc(callback)
    inl = _inloop(callback)
    walls = WallCallback[c for c in inl if c isa WallCallback]
    counters = [c for c in inl if c isa CollisionCounter]
    length(walls) <= 4 || error("at most 4 WallCallbacks (PTL_MAX_WALLS)")
    wd = ntuple(i -> i <= length(walls) ? walldesc(walls[i]) : NOWALL, 4)
    return CallbackDesc(length(walls), isempty(counters) ? 0 : 1, wd), walls, (isempty(counters) ? nothing : first(counters))
end
