# GENERATED from INTEGRATION.md (section 3: the package extension) by scripts/extract_julia_binding.py -- edit the document, not this file.
# Binding of libintegrals_b200.so (C ABI: include/integrals_b200.h) into Integrals.jl; untested here (no Julia in the build image).

module IntegralsB200Ext
using Integrals, SciMLBase, QuadGK
using Integrals: IntegralCache, VectorOf, SampledIntegralCache, ChangeOfVariables, AbstractB200Algorithm,
                 QuadGKB200, HCubatureB200, QuadratureRuleB200, GaussLegendreB200, TrapezoidalB200, SimpsonsB200,
                 transformation_if_inf, transformation_tan_inf, transformation_cot_inf, family_id
const lib = "libintegrals_b200"          # on the loader path, built by `make -C integrals.jl_b200/csrc`

mutable struct B200Context               # cache.cacheval: owns the device scratch across solves
    h::Ptr{Cvoid}
    function B200Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ib200_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r)
        rc == 0 || error(unsafe_string(ccall((:ib200_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        finalizer(c -> ccall((:ib200_destroy, lib), Cvoid, (Ptr{Cvoid},), c.h), new(r[]))
    end
end
check(ctx, rc) = rc == 0 ? nothing :
    (msg = unsafe_string(ccall((:ib200_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.h));
     rc == -1 ? throw(ArgumentError(msg)) : error(msg))

# per-problem status -> retcode.  The reference always reports Success (src/Integrals.jl:338,396) because QuadGK / HCubature only
# return once the tolerance or the caller's maxiters is met.  The engine can also stop because its box / segment store is full;
# that is NOT a successful solve under the reference's contract, so it is reported as MaxIters.  A non-finite estimate throws
# like QuadGK's DomainError for the 1-D solver and is returned as it is for HCubature (which stops and returns, too).
function retcode_of(alg, st::AbstractVector{Int32}, ne::AbstractVector{Int64}, maxiters; slack = 0)
    alg isa QuadGKB200 && any(==(Int32(2)), st) &&
        throw(DomainError(findfirst(==(Int32(2)), st), "integrand produced a non-finite value (problem index given)"))
    capacity = any(i -> st[i] == 1 && ne[i] + slack < maxiters, eachindex(st))      # slack: budgets counted in whole rule applications
    return capacity ? ReturnCode.MaxIters : ReturnCode.Success
end

transform_id(::typeof(transformation_if_inf)) = Int32(0)
transform_id(::typeof(transformation_tan_inf)) = Int32(1)
transform_id(::typeof(transformation_cot_inf)) = Int32(2)
transform_id(f) = throw(ArgumentError("B200 algorithms support the three built-in transformations only"))

# --- intercept ABOVE the closure-building generic method (src/Integrals.jl:186-224): the engine gets the
#     ORIGINAL domain and applies u2t / t2ujac on the device
Integrals.init_cacheval(alg::ChangeOfVariables{F, <:AbstractB200Algorithm}, prob::IntegralProblem) where {F} = B200Context()
function Integrals.__solve(cache::IntegralCache, alg::ChangeOfVariables{F, <:AbstractB200Algorithm},
                           sensealg, udomain, p; kwargs...) where {F}
    Integrals.__solvebp_call(cache, alg.alg, sensealg, udomain, p; transform = transform_id(alg.fu2gv), kwargs...)
end

# batch convention: p :: Vector (one problem) or Matrix nparam x nprob; bounds scalar / Vector(d) / Matrix d x nprob
_mat(p::AbstractVector) = reshape(collect(Float64, p), :, 1)
_mat(p::AbstractMatrix) = Matrix{Float64}(p)
_mat(p::Number) = fill(Float64(p), 1, 1)
_bounds(lb, ub) = (collect(Float64, vec(collect(lb))), collect(Float64, vec(collect(ub))), lb isa AbstractMatrix ? Int32(1) : Int32(0), lb isa AbstractMatrix ? size(lb, 1) : length(lb))

function Integrals.__solvebp_call(cache::IntegralCache, alg::QuadGKB200, sensealg, domain, p;
                                  transform = Int32(0), reltol = 1.0e-8, abstol = 1.0e-8, maxiters = typemax(Int))
    cache.f isa BatchIntegralFunction && return solve_batchfn(cache, alg, domain, p; transform, reltol, abstol, maxiters)   # src/Integrals.jl:274
    lb, ub, bpp, dim = _bounds(domain...)
    dim == 1 || throw(ArgumentError("QuadGKB200 only accepts one-dimensional quadrature problems."))
    cache.f.f isa VectorOf && return solve_vector(cache, alg, lb, ub, bpp, dim, p; transform, reltol, abstol, maxiters)      # src/Integrals.jl:316-326
    P = _mat(p); nprob = size(P, 2)
    I = Vector{Float64}(undef, nprob); E = similar(I); ne = Vector{Int64}(undef, nprob); st = Vector{Int32}(undef, nprob)
    ctx = cache.cacheval
    if alg.order == 7
        GC.@preserve lb ub P I E ne st check(ctx, ccall((:ib200_quadgk_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Float64, Float64, Int64,
             Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
            ctx.h, family_id(cache.f.f), transform, lb, ub, bpp, P, size(P, 1), nprob, reltol, abstol, min(maxiters, 2^62), I, E, ne, st))
    else                                                   # order = alg.order (src/Integrals.jl:331-334): QuadGK's own tables, as they are
        x, w, wg = QuadGK.kronrod(alg.order)
        GC.@preserve lb ub P I E ne st x w wg check(ctx, ccall((:ib200_quadgk_rule_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Float64, Float64, Int64,
             Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
            ctx.h, family_id(cache.f.f), transform, lb, ub, bpp, P, size(P, 1), nprob, reltol, abstol, min(maxiters, 2^62),
            alg.order, x, w, wg, I, E, ne, st))
    end
    prob = Integrals.build_problem(cache)
    u, r = p isa AbstractMatrix ? (I, E) : (I[1], E[1])
    SciMLBase.build_solution(prob, alg, u, r, retcode = retcode_of(alg, st, ne, maxiters))     # Success like src/Integrals.jl:338
end

function Integrals.__solvebp_call(cache::IntegralCache, alg::HCubatureB200, sensealg, domain, p;
                                  transform = Int32(0), reltol = 1.0e-8, abstol = 1.0e-8, maxiters = typemax(Int))
    cache.f isa BatchIntegralFunction && return solve_batchfn(cache, alg, domain, p; transform, reltol, abstol, maxiters)
    alg.single && return solve_single(cache, alg, domain, p; transform, reltol, abstol, maxiters)
    lb, ub, bpp, dim = _bounds(domain...)
    cache.f.f isa VectorOf && return solve_vector(cache, alg, lb, ub, bpp, dim, p; transform, reltol, abstol, maxiters)      # norm = alg.norm, src/Integrals.jl:380-393
    P = _mat(p); nprob = size(P, 2)
    I = Vector{Float64}(undef, nprob); E = similar(I); ne = Vector{Int64}(undef, nprob); st = Vector{Int32}(undef, nprob)
    ctx = cache.cacheval
    rounds = alg.refinement === :rounds || (alg.refinement === :auto && 2 <= dim <= 8)      # IB200_HCUB_ROUNDS is built for 2 <= dim <= 8
    GC.@preserve lb ub P I E ne st check(ctx, ccall((:ib200_hcubature_batch, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Float64, Float64, Int64, Int32, Int32,
         Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
        ctx.h, family_id(cache.f.f), dim, transform, lb, ub, bpp, P, size(P, 1), nprob, reltol, abstol, min(maxiters, 2^62),
        alg.initdiv, rounds ? 1 : 0, I, E, ne, st))
    prob = Integrals.build_problem(cache)
    u, r = p isa AbstractMatrix ? (I, E) : (I[1], E[1])
    SciMLBase.build_solution(prob, alg, u, r, retcode = retcode_of(alg, st, ne, maxiters))     # Success like src/Integrals.jl:396
end

# vector-valued integrands (VectorOf): p :: nparam x ncomp (one problem) or nparam x ncomp x nprob; the components share the nodes and
# the subdivision, the error is a norm over them (QuadGKJL.norm / HCubatureJL.norm)
function solve_vector(cache::IntegralCache, alg::Union{QuadGKB200, HCubatureB200}, lb, ub, bpp, dim, p; transform, reltol, abstol, maxiters)
    P = p isa AbstractMatrix ? reshape(Array{Float64}(p), size(p, 1), size(p, 2), 1) : Array{Float64, 3}(p)
    nparam, ncomp, nprob = size(P)
    I = Matrix{Float64}(undef, ncomp, nprob); E = Vector{Float64}(undef, nprob); ne = Vector{Int64}(undef, nprob); st = Vector{Int32}(undef, nprob)
    ctx = cache.cacheval
    normkind = Int32(alg.norm === :inf ? 1 : 0)
    if alg isa QuadGKB200
        GC.@preserve lb ub P I E ne st check(ctx, ccall((:ib200_quadgk_vec_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int32, Int64, Float64, Float64, Int64, Int32,
             Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
            ctx.h, family_id(cache.f.f), transform, lb, ub, bpp, P, nparam, ncomp, nprob, reltol, abstol, min(maxiters, 2^62), normkind, I, E, ne, st))
    else
        GC.@preserve lb ub P I E ne st check(ctx, ccall((:ib200_hcubature_vec_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int32, Int64, Float64, Float64, Int64, Int32, Int32,
             Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
            ctx.h, family_id(cache.f.f), dim, transform, lb, ub, bpp, P, nparam, ncomp, nprob, reltol, abstol, min(maxiters, 2^62), alg.initdiv, normkind,
            I, E, ne, st))
    end
    u, r = p isa AbstractMatrix ? (I[:, 1], E[1]) : (I, E)
    SciMLBase.build_solution(Integrals.build_problem(cache), alg, u, r, retcode = retcode_of(alg, st, ne, maxiters))
end

# parameter sensitivities: I and dI/dp[wrt] of every problem as ONE vector-valued integral [f, df/dp...] on shared nodes and subdivision.
# The engine cannot push ForwardDiff.Dual numbers or a Zygote pullback through a device-side integrand, so the AD rules of the reference
# (ext/IntegralsForwardDiffExt.jl:55-134, ext/IntegralsZygoteExt.jl:144-178) are replaced for the registered families by this call:
#   a `__solvebp(cache, alg::AbstractB200Algorithm, sensealg, domain, p::AbstractArray{<:ForwardDiff.Dual})` method calls it with
#   control = :value (dual numbers flowing through QuadGKJL / HCubatureJL leave the error control to the value) and rebuilds the duals
#   from (I, dI/dp * partials); the ChainRules `rrule` calls it with control = :norm2 and returns dI/dp' * ybar.
# wrt: 1-based parameter rows (at most 7 per call; more are solved in groups); control: :value, :norm2 or :norminf.
function b200_sensitivities(prob::IntegralProblem, alg::Union{QuadGKB200, HCubatureB200}; wrt = nothing, control::Symbol = :value,
                            reltol = 1.0e-8, abstol = 1.0e-8, maxiters = typemax(Int), transform::Int32 = Int32(0))
    cache = init(prob, alg)
    ctx = cache.cacheval
    lb0, ub0 = prob.domain
    dim = length(lb0)
    P = prob.p isa AbstractMatrix ? Matrix{Float64}(prob.p) : reshape(collect(Float64, prob.p), :, 1)
    nparam, nprob = size(P)
    rows = wrt === nothing ? collect(1:nparam) : collect(Int, wrt)
    lb = collect(Float64, vec(collect(lb0))); ub = collect(Float64, vec(collect(ub0)))
    normkind = Int32(control === :value ? 2 : (control === :norminf ? 1 : 0))
    Iv = Vector{Float64}(undef, nprob); dI = Matrix{Float64}(undef, length(rows), nprob)
    E = zeros(nprob); ne = zeros(Int64, nprob); st = zeros(Int32, nprob)
    for g0 in 1:7:length(rows)
        grp = rows[g0:min(g0 + 6, length(rows))]
        idx = Int32.(grp .- 1); ns = Int32(length(idx))
        out = Matrix{Float64}(undef, 1 + ns, nprob); Eg = similar(E); neg = similar(ne); stg = similar(st)
        if alg isa QuadGKB200
            GC.@preserve lb ub P idx out Eg neg stg check(ctx, ccall((:ib200_quadgk_sens_batch, lib), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Ptr{Int32}, Int32, Float64, Float64, Int64, Int32,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
                ctx.h, family_id(prob.f.f), transform, lb, ub, Int32(0), P, nparam, nprob, idx, ns, reltol, abstol, min(maxiters, 2^62), normkind,
                out, Eg, neg, stg))
        else
            GC.@preserve lb ub P idx out Eg neg stg check(ctx, ccall((:ib200_hcubature_sens_batch, lib), Int32,
                (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Ptr{Int32}, Int32, Float64, Float64, Int64,
                 Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int32}),
                ctx.h, family_id(prob.f.f), dim, transform, lb, ub, Int32(0), P, nparam, nprob, idx, ns, reltol, abstol, min(maxiters, 2^62),
                alg.initdiv, normkind, out, Eg, neg, stg))
        end
        g0 == 1 && (Iv .= out[1, :])
        dI[g0:(g0 + ns - 1), :] .= out[2:end, :]
        E .= max.(E, Eg); ne .= max.(ne, neg); st .= max.(st, stg)
    end
    u, r = prob.p isa AbstractMatrix ? (Iv, E) : (Iv[1], E[1])
    sol = SciMLBase.build_solution(prob, alg, u, r, retcode = retcode_of(alg, st, ne, maxiters))
    return sol, (prob.p isa AbstractMatrix ? dI : dI[:, 1])
end

# BatchIntegralFunction (src/Integrals.jl:274-315): ANY Julia closure f(pts, p) -> values, evaluated by Julia on the batches of
# points the engine hands out; the points of a round, the weighted sums, the error estimates and the refinement stay on the
# device.  With CuArray buffers (pointer(x::CuArray)) nothing is copied; a host Array works the same way.
function solve_batchfn(cache::IntegralCache, alg::Union{QuadGKB200, HCubatureB200}, domain, p;
                       transform = Int32(0), reltol = 1.0e-8, abstol = 1.0e-8, maxiters = typemax(Int))
    f = cache.f
    lb, ub, _, dim = _bounds(domain...)
    alg isa QuadGKB200 && dim != 1 && throw(ArgumentError("QuadGKB200 only accepts one-dimensional quadrature problems."))
    npts = dim == 1 ? 15 : 1 + 4dim + 2dim * (dim - 1) + 2^dim
    ctx = cache.cacheval
    GC.@preserve lb ub check(ctx, ccall((:ib200_batch_open, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int64, Int64),
        ctx.h, dim, transform, lb, ub, reltol, abstol, max(1, min(maxiters, 2^62) ÷ npts), 0))
    n = Ref{Int64}(0)
    while true
        check(ctx, ccall((:ib200_batch_next, lib), Int32, (Ptr{Cvoid}, Ref{Int64}), ctx.h, n))
        n[] == 0 && break
        x = dim == 1 ? Vector{Float64}(undef, n[]) : Matrix{Float64}(undef, dim, n[])         # or CuArray{Float64}(undef, dim, n[])
        GC.@preserve x check(ctx, ccall((:ib200_batch_points, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx.h, pointer(x), n[]))
        y = isinplace(f) ? (yy = similar(x, n[]); f(yy, x, p); yy) : f(x, p)                   # the reference's two call styles
        y = convert(typeof(similar(x, n[])), y)
        GC.@preserve y check(ctx, ccall((:ib200_batch_values, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx.h, pointer(y), n[]))
    end
    I = Ref(0.0); E = Ref(0.0); stats = zeros(Int64, 6)
    GC.@preserve stats check(ctx, ccall((:ib200_batch_result, lib), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ptr{Int64}), ctx.h, I, E, stats))
    SciMLBase.build_solution(Integrals.build_problem(cache), alg, I[], E[],
                             retcode = retcode_of(alg, Int32[stats[5]], Int64[stats[1]], maxiters; slack = 2npts))
end
# (device y must be complete -- CUDA.synchronize() -- before ib200_batch_values; f.max_batch: split the round's points into views of at most
#  max_batch columns and call f on each, as the Python mirror does -- the engine does not care how y was produced)

# one expensive integral (BASELINE config 5): HCubatureB200(single = true).  `maxiters` bounds the integrand evaluations
# of the whole job, as for HCubatureJL; the engine's budget is in rule applications (boxes).  With several GPUs every
# rank (one Julia process per GPU, e.g. under MPI.jl) makes this same call; see "Multi-GPU" below.
function solve_single(cache::IntegralCache, alg::HCubatureB200, domain, p; transform, reltol, abstol, maxiters)
    lb, ub, _, dim = _bounds(domain...)
    npts = 1 + 4dim + 2dim * (dim - 1) + 2^dim
    I = Ref(0.0); E = Ref(0.0); stats = zeros(Int64, 8); P = collect(Float64, vec(p)); ctx = cache.cacheval
    GC.@preserve lb ub P stats check(ctx, ccall((:ib200_hcubature_single, lib), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Float64, Int64, Int64,
         Ref{Float64}, Ref{Float64}, Ptr{Int64}),
        ctx.h, family_id(cache.f.f), dim, transform, lb, ub, P, length(P), reltol, abstol, max(1, min(maxiters, 2^62) ÷ npts), 0, I, E, stats))
    SciMLBase.build_solution(Integrals.build_problem(cache), alg, I[], E[],
                             retcode = retcode_of(alg, Int32[stats[5]], Int64[stats[1]], maxiters; slack = 2npts))
end

function Integrals.__solvebp_call(cache::IntegralCache, alg::QuadratureRuleB200, sensealg, domain, p;
                                  transform = Int32(0), reltol = nothing, abstol = nothing, maxiters = nothing)
    lb, ub, bpp, dim = _bounds(domain...)
    P = _mat(p); nprob = size(P, 2); I = Vector{Float64}(undef, nprob)
    x, w = alg.q(alg.n); ctx = cache.cacheval
    if alg.tensor
        xs = collect(Float64, x); ws = collect(Float64, w)
        GC.@preserve lb ub P I xs ws check(ctx, ccall((:ib200_fixed_rule_tensor_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}),
            ctx.h, family_id(cache.f.f), dim, transform, lb, ub, bpp, P, size(P, 1), nprob, xs, ws, length(xs), I))
    else
        nodes = Matrix{Float64}(reduce(hcat, collect.(x)))          # dim x N column-major == N x dim row-major
        ws = collect(Float64, w)
        GC.@preserve lb ub P I nodes ws check(ctx, ccall((:ib200_fixed_rule_nodes_batch, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
            ctx.h, family_id(cache.f.f), dim, transform, lb, ub, bpp, P, size(P, 1), nprob, nodes, ws, length(ws), I))
    end
    SciMLBase.build_solution(Integrals.build_problem(cache), alg, p isa AbstractMatrix ? I : I[1], nothing, retcode = ReturnCode.Success)
end

# sampled (src/sampled.jl:127-168).  `init` calls init_cacheval(alg, prob) = find_weights(prob.x, alg) (src/sampled.jl:128-133) and
# __solvebp_call recomputes cache.cacheval = find_weights(cache.x, alg) when the grid was replaced (cache.isfresh, src/common.jl:276-281).
# The engine forms the weights on the device from the grid (uniform: from (n, h) alone), so the "weights" object of the B200 rules is
# the handle of that grid: the context + the step of an AbstractRange, or a Float64 copy of the points that stays rooted here.
struct B200SampledWeights
    ctx::B200Context
    h::Float64                 # step(x) for an AbstractRange (uniform weights, src/trapezoidal.jl:43, src/simpsons.jl:85)
    xs::Vector{Float64}        # the grid points otherwise (empty for a range)
end
const default_ctx = Ref{Union{Nothing, B200Context}}(nothing)
default_context() = (default_ctx[] === nothing && (default_ctx[] = B200Context()); default_ctx[]::B200Context)
function Integrals.find_weights(x::AbstractVector, alg::Union{TrapezoidalB200, SimpsonsB200})
    alg isa SimpsonsB200 && !(x isa AbstractRange) && length(x) <= 2 &&
        throw(ArgumentError("The length of the grid must exceed 2 for simpsons rule."))          # src/simpsons.jl:46
    x isa AbstractRange && return B200SampledWeights(default_context(), Float64(step(x)), Float64[])
    return B200SampledWeights(default_context(), 0.0, collect(Float64, x))
end
# host Array: its own pointer.  Device arrays are passed by address too (the C ABI resolves host / device per pointer); CUDA.jl's
# pointer(::CuArray) is a CuPtr, converted through its integer value.  CUDA.jl runs on per-task streams, so the engine joins the
# caller's stream first (ib200_set_stream) -- see IntegralsB200CUDAExt below; without it only host arrays are accepted.
dataptr(a::Array{Float64}) = pointer(a)
dataptr(a) = throw(ArgumentError("TrapezoidalB200 / SimpsonsB200: y must be an Array{Float64} (or a CuArray{Float64} with CUDA.jl loaded); got $(typeof(a))"))
join_stream(ctx, a) = nothing
function Integrals.__solvebp_call(cache::SampledIntegralCache, alg::Union{TrapezoidalB200, SimpsonsB200}; kwargs...)
    if cache.isfresh                                                   # like src/sampled.jl:144-151
        cache.cacheval = Integrals.find_weights(cache.x, alg)
        cache.isfresh = false
    end
    w = cache.cacheval::B200SampledWeights
    y, dim = cache.y, Integrals.dimension(cache.dim)
    n = size(y, dim)
    n == 0 && error("No points to integrate")                          # src/sampled.jl:55
    inner = prod(size(y)[1:(dim - 1)]); outer = prod(size(y)[(dim + 1):end])
    out = similar(y, Float64, (size(y)[1:(dim - 1)]..., size(y)[(dim + 1):end]...))
    xs = w.xs                                                          # a local binding: rooted by GC.@preserve for the whole ccall
    xp = isempty(xs) ? Ptr{Float64}(C_NULL) : pointer(xs)
    join_stream(w.ctx, y)
    GC.@preserve y out xs check(w.ctx, ccall((:ib200_sampled, lib), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}),
        w.ctx.h, alg isa SimpsonsB200 ? 1 : 0, dataptr(y), inner, n, outer, w.h, xp, dataptr(out)))
    check(w.ctx, ccall((:ib200_synchronize, lib), Int32, (Ptr{Cvoid},), w.ctx.h))     # device outputs are asynchronous on the engine's stream
    I = ndims(out) == 0 ? out[] : out
    SciMLBase.build_solution(Integrals.build_problem(cache), alg, I, nothing, retcode = ReturnCode.Success)   # src/sampled.jl:166-167, src/common.jl:412
end
end # module
