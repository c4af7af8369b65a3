# GENERATED from INTEGRATION.md (section 2: algorithm structs and registered integrands, next to src/algorithms_extension.jl) by scripts/extract_julia_binding.py -- edit the document, not this file.
# Binding of libintegrals_b200.so (C ABI: include/integrals_b200.h) into Integrals.jl; untested here (no Julia in the build image).

abstract type AbstractB200Algorithm <: AbstractIntegralCExtensionAlgorithm end

struct QuadGKB200 <: AbstractB200Algorithm                          # like QuadGKJL(; order = 7, norm = norm)
    order::Int              # 7: built-in tables; 2..30: QuadGK.kronrod(order) is passed to the engine
    norm::Symbol            # :two (LinearAlgebra.norm) or :inf -- only used by vector-valued integrands (VectorOf)
end
QuadGKB200(; order = 7, norm = :two) = QuadGKB200(order, norm)
struct HCubatureB200 <: AbstractB200Algorithm                       # like HCubatureJL(; initdiv = 1, norm = norm)
    initdiv::Int
    single::Bool            # true: one domain region-partitioned over all GPUs
    norm::Symbol            # :two or :inf -- only used by vector-valued integrands (VectorOf)
    refinement::Symbol      # :rounds (every box above a per-problem threshold per round; 2 <= dim <= 8), :reference_order (one box
end                         # per step, HCubature's own sequence of boxes) or :auto (rounds where they are built)
HCubatureB200(; initdiv = 1, single = false, norm = :two, refinement = :auto) = HCubatureB200(initdiv, single, norm, refinement)
struct QuadratureRuleB200{Q} <: AbstractB200Algorithm
    q::Q; n::Int; tensor::Bool
end
QuadratureRuleB200(q; n = 250, tensor = false) = QuadratureRuleB200(q, n, tensor)
struct GaussLegendreB200 <: AbstractB200Algorithm
    nodes::Vector{Float64}; weights::Vector{Float64}; subintervals::Int
end
struct TrapezoidalB200 <: AbstractSampledIntegralAlgorithm end
struct SimpsonsB200 <: AbstractSampledIntegralAlgorithm end

# registered device-side integrands: callable structs, so the SAME problem also runs under
# QuadGKJL / HCubatureJL for parity checks
struct SinXP end;            (::SinXP)(x, p) = sin(x * p[1])
struct GenzOscillatory end;  (::GenzOscillatory)(x, p) = (d = length(x); cos(2π * p[d + 1] + sum(p[i] * x[i] for i in 1:d)))
struct GenzProductPeak end;  (::GenzProductPeak)(x, p) = (d = length(x); 1 / prod(p[i]^-2 + (x[i] - p[d + i])^2 for i in 1:d))
struct GenzCornerPeak end;   (::GenzCornerPeak)(x, p) = (d = length(x); (1 + sum(p[i] * x[i] for i in 1:d))^(-(d + 1)))
struct GenzGaussian end;     (::GenzGaussian)(x, p) = (d = length(x); exp(-sum((p[i] * (x[i] - p[d + i]))^2 for i in 1:d)))
struct GenzC0 end;           (::GenzC0)(x, p) = (d = length(x); exp(-sum(p[i] * abs(x[i] - p[d + i]) for i in 1:d)))
struct GenzDiscontinuous end
(::GenzDiscontinuous)(x, p) = (d = length(x); (x[1] > p[d + 1] || (d > 1 && x[2] > p[d + 2])) ? 0.0 : exp(sum(p[i] * x[i] for i in 1:d)))
family_id(::SinXP) = 0; family_id(::GenzOscillatory) = 1; family_id(::GenzProductPeak) = 2
family_id(::GenzCornerPeak) = 3; family_id(::GenzGaussian) = 4; family_id(::GenzC0) = 5; family_id(::GenzDiscontinuous) = 6
# vector-valued integrand built from one registered family: component c uses the parameter column p[:, c]
struct VectorOf{F} base::F end; (v::VectorOf)(x, p) = [v.base(x, view(p, :, c)) for c in axes(p, 2)]
family_id(v::VectorOf) = family_id(v.base)
family_id(f) = throw(ArgumentError("the B200 back-end integrates registered integrands only; got $(typeof(f))"))
