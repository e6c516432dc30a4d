# tables.jl — host-built collision tables -> device tables.  Table CONSTRUCTION (chebfit, totalcs, the Seltzer-Berger
# inverse CDF, load_lxcat) is init code in the reference and stays in Particulator.jl; only the flat arrays travel.

procdesc(ctx, ::NullCollision)        = ProcessDesc(PROC_NULL, -1, ntuple(_ -> 0.0, 6))
procdesc(ctx, p::RelativisticCoulomb) = ProcessDesc(PROC_COULOMB, -1, (Float64(p.Z), 0.0, 0.0, 0.0, 0.0, 0.0))        # relativistic_coulomb.jl:5-8
procdesc(ctx, p::RBEB)                = ProcessDesc(PROC_RBEB, -1, (Float64(p.B), Float64(p.U), Float64(p.N), 0.0, 0.0, 0.0))   # rbeb.jl:5-17
procdesc(ctx, p::Moller)              = ProcessDesc(PROC_MOLLER, -1, (Float64(p.Z), Float64(p.tcut), 0.0, 0.0, 0.0, 0.0))       # moller.jl:8-11
procdesc(ctx, p::Bhaba)               = ProcessDesc(PROC_BHABA, -1, (Float64(p.Z), Float64(p.tcut), 0.0, 0.0, 0.0, 0.0))        # bhaba.jl:4-7
procdesc(ctx, p::Union{Compton,KleinNishinaCompton}) = ProcessDesc(PROC_COMPTON, -1, (Float64(p.Z), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p::BetheHeitler)        = ProcessDesc(PROC_BETHE_HEITLER, -1, (Float64(p.Z), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p::PositronAnihilation) = ProcessDesc(PROC_ANIHILATION, -1, (Float64(p.Z), 0.0, 0.0, 0.0, 0.0, 0.0))
function procdesc(ctx, p::PhotoElectric)                                   # photo_electric.jl:7-32: shells scanned K-first
    b = p.binding
    nb = min(4, length(b))
    ProcessDesc(PROC_PHOTOELECTRIC, -1, (Float64(p.Z), Float64(nb), ntuple(i -> i <= nb ? Float64(b[i]) : 0.0, 4)...))
end
function procdesc(ctx, sb::SeltzerBerger)                                  # seltzer.jl:9-49: data[ncum, nE], ncum fastest
    id = get!(ctx.sb, sb) do
        data = Matrix{Float64}(sb.data)
        le = Vector{Float64}(sb.log_energy)
        rc = GC.@preserve data le ccall((:ptl_sb_table_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}),
                                        ctx.h, size(data, 1), size(data, 2), le, data)
        rc < 0 && error("ptl_sb_table_create: status $rc: $(last_error(ctx))")
        rc
    end
    ProcessDesc(PROC_SELTZER, id, ntuple(_ -> 0.0, 6))
end
# LXCat kinds (slow-electron.jl:70-84)
procdesc(ctx, p::Particulator.Excitation) = ProcessDesc(PROC_LX_EXCITATION, -1, (Float64(p.threshold), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p::Particulator.Ionization) = ProcessDesc(PROC_LX_IONIZATION, -1, (Float64(p.threshold), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p::Particulator.Attachment) = ProcessDesc(PROC_LX_ATTACHMENT, -1, (Float64(p.threshold), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p::Particulator.Elastic)    = ProcessDesc(PROC_LX_ELASTIC, -1, (Float64(p.mass_ratio), 0.0, 0.0, 0.0, 0.0, 0.0))
procdesc(ctx, p) = error("process $(typeof(p)) has no device implementation (PhotoEmission and Zhelezniak photons are out of scope)")

"ChebyshevCollisionTable (collision_table.jl:63-75): rate[order, nprocs, k+1], ratebound[order, k+1], column-major as stored."
function device_table(ctx::Context, c::ChebyshevCollisionTable{T,N}) where {T,N}
    get!(ctx.tables, c) do
        procs = ProcessDesc[procdesc(ctx, p) for p in c.proc]
        rate = Array{Float64,3}(c.rate)
        rb = Matrix{Float64}(c.ratebound)                    # chebfit returns [order, k+1] (cheby.jl:211-227)
        rc = GC.@preserve procs rate rb ccall((:ptl_table_create_cheb, LIB), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{ProcessDesc}),
            ctx.h, N, length(procs), c.b.k, Float64(c.b.xmax), rate, rb, procs)
        rc < 0 && error("ptl_table_create_cheb: status $rc: $(last_error(ctx))")
        rc
    end
end

"CollisionTable (collision_table.jl:15-57) on a LinRange or LogLinRange grid; scalar or vector rate bound (:33-43)."
function device_table(ctx::Context, c::CollisionTable)
    get!(ctx.tables, c) do
        procs = ProcessDesc[procdesc(ctx, p) for p in c.proc]
        rate = Matrix{Float64}(c.rate)                       # [nprocs, nE], process fastest
        e = c.energy
        if e isa LogLinRange
            kind, L1, L2 = Int32(1), Float64(first(e.L)), Float64(last(e.L))        # util.jl:60-77: x = exp(L) - exp(L[1])
        else
            kind, L1, L2 = Int32(0), Float64(first(e)), Float64(last(e))
        end
        rc = if c.ratebound isa Number
            GC.@preserve procs rate ccall((:ptl_table_create_linear, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Float64, Ptr{ProcessDesc}),
                ctx.h, kind, L1, L2, length(e), length(procs), rate, Float64(c.ratebound), procs)
        else
            rb = Vector{Float64}(c.ratebound)
            GC.@preserve procs rate rb ccall((:ptl_table_create_linear_vb, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Float64, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{ProcessDesc}),
                ctx.h, kind, L1, L2, length(e), length(procs), rate, rb, procs)
        end
        rc < 0 && error("ptl_table_create_linear: status $rc: $(last_error(ctx))")
        rc
    end
end

"ChebContinuumLoss{N} (continuum.jl:25-43): coefficient matrices ec, pc [N, k+1]."
function device_cheb_loss(ctx::Context, cl::ChebContinuumLoss{N}) where N
    get!(ctx.cheb_losses, cl) do
        ec, pc = Matrix{Float64}(cl.ec), Matrix{Float64}(cl.pc)
        rc = GC.@preserve ec pc ccall((:ptl_cheb_loss_create, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Ptr{Float64}, Ptr{Float64}),
                                      ctx.h, N, cl.bints.k, Float64(cl.bints.xmax), ec, pc)
        rc < 0 && error("ptl_cheb_loss_create: status $rc: $(last_error(ctx))")
        rc
    end
end

"rate(table, j, presample(E)) for every process and ratebound(E), evaluated by the device functions the advance kernel uses (bit-exact tier)."
function table_eval(ctx::Context, table, energy::AbstractVector{<:Real})
    tid = device_table(ctx, table)
    e = Vector{Float64}(energy)
    rates = Matrix{Float64}(undef, length(table.proc), length(e))
    bound = Vector{Float64}(undef, length(e))
    rc = GC.@preserve e rates bound ccall((:ptl_table_eval, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                          ctx.h, tid, length(e), e, rates, bound)
    check(ctx, rc, "table_eval")
    return rates, bound
end
