# comm.jl — multi-GPU: one Julia process (or task) per GPU, each with its own Context and shard.  The advance path needs no
# collective; NCCL (called from inside the library) reduces what run! prints / the population-control callbacks decide on
# (run.jl:31-40, callback.jl:203,217,239,263) and rebalances the populations.

"`ncclGetUniqueId`: call on ONE rank and ship the 128 bytes to the others (MPI.Bcast!, Distributed, a file, ...)."
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:ptl_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
    rc == 0 || error("ptl_comm_unique_id: status $rc (libnccl.so.2 not found? set PTL_NCCL_LIB)")
    return id
end
comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ctx, GC.@preserve(id, ccall((:ptl_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), ctx.h, id, rank, nranks)), "comm_init!")
comm_destroy!(ctx::Context) = check(ctx, ccall((:ptl_comm_destroy, LIB), Int32, (Ptr{Cvoid},), ctx.h), "comm_destroy!")

"Diagnostics with GLOBAL sums / max over all ranks (what `_msg` prints, run.jl:31-40)."
function diag_allreduce(d::DevicePopulation)
    o = Ref{DiagOut}()
    check(d.ctx, ccall((:ptl_diag_allreduce, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{DiagOut}), d.ctx.h, d.id, o), "diag_allreduce")
    return o[]
end
function histogram_allreduce(d::DevicePopulation, quantity::Symbol, lo, hi, nbins; logscale = false)
    out = zeros(Float64, nbins)
    q = quantity === :energy ? 0 : 1
    check(d.ctx, GC.@preserve(out, ccall((:ptl_histogram_allreduce, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Float64, Int32, Int32, _PF),
                                         d.ctx.h, d.id, q, Float64(lo), Float64(hi), nbins, logscale ? 1 : 0, out)), "histogram_allreduce")
    return out
end
"In-place all-reduce of a small host vector (`op` = :sum, :max, :min): global counts for RouletteCallback / PopulationTargetCallback."
function allreduce!(ctx::Context, v::Vector{Float64}, op::Symbol = :sum)
    o = op === :sum ? 0 : op === :max ? 1 : 2
    check(ctx, GC.@preserve(v, ccall((:ptl_comm_allreduce_f64, LIB), Int32, (Ptr{Cvoid}, _PF, Int32, Int32), ctx.h, v, length(v), o)), "allreduce!")
    return v
end
"Rebalance one species over the communicator (collective). Returns (n after, rows sent (+) or received (-))."
function rebalance!(d::DevicePopulation; tolerance = 0.05)
    moved = Ref{Int64}(0)
    n = ccall((:ptl_rebalance, LIB), Int64, (Ptr{Cvoid}, Int32, Float64, Ref{Int64}), d.ctx.h, d.id, Float64(tolerance), moved)
    n < 0 && error("rebalance!: status $n: $(last_error(d.ctx))")
    return Int(n), Int(moved[])
end
