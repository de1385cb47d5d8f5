# SthenoB200.jl -- thin `ccall` shim that routes Stheno/AbstractGPs' dense-GP hot path to
# libstheno_b200.so (include/stheno_b200.h).
#
# STATUS: written against Stheno v0.8.2 / AbstractGPs 0.5 / KernelFunctions 0.10; it has NOT been
# executed (no `julia` binary in the build image or on the GPU box).  The Python mirror
# stheno.jl_b200/{gp,finite,lib}.py binds the very same symbols call for call and is what the
# tests and the benchmark run; keep the two in sync.
#
# Usage
#     using Stheno, SthenoB200
#     f  = @gppp let f1 = GP(SEKernel()); f2 = GP(Matern52Kernel()); f3 = f1 + f2 end
#     fb = b200(f)                                  # device-marked programme
#     fx = fb(BlockData(GPPPInput(:f1, x1), GPPPInput(:f3, x3)), 0.1)
#     logpdf(fx, y); fp = posterior(fx, y); mean_and_var(fp(GPPPInput(:f2, xs)))
module SthenoB200

using Stheno, AbstractGPs, KernelFunctions, LinearAlgebra
using Stheno: AtomicGP, DerivedGP, GPPP, BlockData, GPPPInput, SthenoAbstractGP
import AbstractGPs: logpdf, posterior, mean, var, cov, mean_and_var, marginals, rand, elbo

const LIB = get(ENV, "STHENO_B200_LIB", "libstheno_b200.so")

# ---- C structs (layout == include/stheno_b200.h) ---------------------------------------------
struct SbArray;  data::Ptr{Cvoid}; n::Int64; dim::Int32; reserved::Int32; end
struct SbTerm;   kernel::Int32; zl::Int32; zr::Int32; sl::Int32; sr::Int32; reserved::Int32
                 coeff::Float64; param::Float64; end
struct SbBlock;  row0::Int64; nrows::Int64; col0::Int64; ncols::Int64; term0::Int32; nterms::Int32; end
struct SbCovSpec
    nrows::Int64; ncols::Int64; symmetric::Int32; narrays::Int32; arrays::Ptr{SbArray}
    nterms::Int32; terms::Ptr{SbTerm}; nblocks::Int32; blocks::Ptr{SbBlock}
end
struct SbNoise;  sigma2::Float64; diag::Ptr{Cvoid}; dense::Ptr{Cvoid}; end

const K_SE, K_M12, K_M32, K_M52, K_WHITE, K_CONST = Int32.(0:5)

function check(status::Int32, info::Int64=0)
    status == 0 && return
    msg = unsafe_string(ccall((:sb_last_error, LIB), Cstring, ()))
    status == -3 && throw(LinearAlgebra.PosDefException(info))   # what `cholesky` throws
    status == -4 && error("SthenoB200: unsupported: $msg")
    error("SthenoB200 (status $status): $msg")
end

# ---- context / handles ------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sb_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[]); finalizer(c -> ccall((:sb_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.h), c); c
    end
end
const CTX = Ref{Union{Nothing,Context}}(nothing)
ctx() = (CTX[] === nothing && (CTX[] = Context()); CTX[])

mutable struct Factor
    h::Ptr{Cvoid}; n::Int
end
destroy!(F::Factor) = (F.h != C_NULL && ccall((:sb_factor_destroy, LIB), Int32, (Ptr{Cvoid},), F.h); F.h = C_NULL)

# ---- plan lowering (SURVEY App. B.3): process -> [(atom, coeff, scale, z)] ---------------------
struct LTerm; atom::AtomicGP; coeff::Float64; scale::Union{Nothing,Vector{Float64}}; z; end

lower(f::AtomicGP, x) = (mean(f, x), [LTerm(f, 1.0, nothing, x)])
function lower(f::DerivedGP, x)
    op = f.args[1]
    if op === (+) && f.args[2] isa AbstractGPs.AbstractGP          # addition.jl:26-47
        ma, ta = lower(f.args[2], x); mb, tb = lower(f.args[3], x)
        return ma .+ mb, vcat(ta, tb)
    elseif op === (+)                                              # addition.jl:73-86 (known)
        b = f.args[2]; m, t = lower(f.args[3], x)
        return (b isa Real ? b .+ m : b.(x) .+ m), t
    elseif op === (*)                                              # product.jl:25-70
        s = f.args[2]; m, t = lower(f.args[3], x)
        if s isa Real
            return s .* m, [LTerm(u.atom, s * u.coeff, u.scale, u.z) for u in t]
        end
        sx = Float64.(s.(x))
        return sx .* m, [LTerm(u.atom, u.coeff, u.scale === nothing ? sx : sx .* u.scale, u.z) for u in t]
    elseif op === (∘)                                              # compose.jl:16-28
        return lower(f.args[2], f.args[3].(x))
    end
    error("SthenoB200: cannot lower $(op)")
end

# kernel -> [(coeff, id, param, input scale)]
klower(::SEKernel) = [(1.0, K_SE, 0.0, 1.0)]
klower(::Matern12Kernel) = [(1.0, K_M12, 0.0, 1.0)]
klower(::Matern32Kernel) = [(1.0, K_M32, 0.0, 1.0)]
klower(::Matern52Kernel) = [(1.0, K_M52, 0.0, 1.0)]
klower(::WhiteKernel) = [(1.0, K_WHITE, 0.0, 1.0)]
klower(k::ConstantKernel) = [(1.0, K_CONST, Float64(only(k.c)), 1.0)]
klower(k::ScaledKernel) = [(c * only(k.σ²), id, p, s) for (c, id, p, s) in klower(k.kernel)]
klower(k::KernelSum) = reduce(vcat, klower.(k.kernels))
klower(k::TransformedKernel{<:Any,<:ScaleTransform}) =
    [(c, id, p, s * only(k.transform.s)) for (c, id, p, s) in klower(k.kernel)]

pointmajor(z::AbstractVector{<:Real}, s) = reshape(Float64.(z) .* s, 1, :)   # 1 x n  (dim 1)
pointmajor(z::ColVecs, s) = Float64.(z.X) .* s                               # D x n  == point-major

# Builds the ccall-able spec; `keep` holds every Julia array the pointers refer to
# (GC.@preserve keep ... around the ccall).
function build_spec(procs_r, xs_r, procs_c, xs_c; symmetric::Bool)
    arrays = SbArray[]; terms = SbTerm[]; blocks = SbBlock[]; keep = Any[]
    function push_arr!(a::Matrix{Float64}, dim)
        push!(keep, a); push!(arrays, SbArray(pointer(a), size(a, 2), dim, 0)); Int32(length(arrays) - 1)
    end
    push_scale!(s) = s === nothing ? Int32(-1) : push_arr!(reshape(s, 1, :), 0)
    lr = [lower(p, x)[2] for (p, x) in zip(procs_r, xs_r)]
    lc = symmetric ? lr : [lower(p, x)[2] for (p, x) in zip(procs_c, xs_c)]
    r0 = cumsum([0; length.(xs_r)]); c0 = cumsum([0; length.(xs_c)])
    for i in eachindex(procs_r), j in eachindex(procs_c)
        symmetric && j > i && continue
        t0 = length(terms)
        for a in lr[i], b in lc[j]
            a.atom === b.atom || continue                            # atomic_gp.jl:36-38
            for (kc, id, p, s) in klower(a.atom.gp.kernel)
                zl = push_arr!(pointmajor(a.z, s), size(pointmajor(a.z, s), 1))
                zr = push_arr!(pointmajor(b.z, s), size(pointmajor(b.z, s), 1))
                push!(terms, SbTerm(id, zl, zr, push_scale!(a.scale), push_scale!(b.scale), 0,
                                    a.coeff * b.coeff * kc, p))
            end
        end
        push!(blocks, SbBlock(r0[i], length(xs_r[i]), c0[j], length(xs_c[j]), t0, length(terms) - t0))
    end
    push!(keep, arrays, terms, blocks)
    spec = SbCovSpec(r0[end], c0[end], symmetric, length(arrays), pointer(arrays), length(terms),
                     pointer(terms), length(blocks), pointer(blocks))
    return spec, keep
end

# ---- device-marked programme and the AbstractGPs methods it overrides ---------------------------
struct B200GPPP{T<:GPPP} <: AbstractGPs.AbstractGP; f::T; end
b200(f::GPPP) = B200GPPP(f)

components(f::B200GPPP, x::GPPPInput) = ([f.f.fs[x.p]], [x.x])
function components(f::B200GPPP, x::BlockData)
    ps = Any[]; vs = Any[]
    for b in x.X; p, v = components(f, b); append!(ps, p); append!(vs, v); end
    ps, vs
end

const B200Finite = AbstractGPs.FiniteGP{<:B200GPPP}
host_mean(fx::B200Finite) = reduce(vcat, [lower(p, x)[1] for (p, x) in zip(components(fx.f, fx.x)...)])

function factor(fx::B200Finite)
    ps, vs = components(fx.f, fx.x)
    spec, keep = build_spec(ps, vs, ps, vs; symmetric=true)
    Σ = fx.Σy
    nd = Σ isa Diagonal ? collect(Float64, diag(Σ)) : Matrix{Float64}(Σ)   # Diagonal or dense PSD Σy
    noise = Σ isa Diagonal ? SbNoise(0.0, pointer(nd), C_NULL) : SbNoise(0.0, C_NULL, pointer(nd))
    h = Ref{Ptr{Cvoid}}(C_NULL); info = Ref{Int64}(0)
    GC.@preserve keep nd begin
        st = ccall((:sb_factor_create, LIB), Int32,
                   (Ptr{Cvoid}, Ref{SbCovSpec}, Ref{SbNoise}, Ref{Ptr{Cvoid}}, Ref{Int64}),
                   ctx().h, spec, noise, h, info)
    end
    check(st, info[])
    F = Factor(h[], length(fx.x)); finalizer(destroy!, F); F
end

function logpdf(fx::B200Finite, y::AbstractVector{<:Real})
    F = factor(fx); δ = Float64.(y .- host_mean(fx)); out = Ref{Float64}(0.0)
    GC.@preserve δ check(ccall((:sb_logpdf, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Ref{Float64}), ctx().h, F.h, δ, 1, out))
    out[]
end

struct B200PosteriorGP{T<:B200GPPP} <: AbstractGPs.AbstractGP
    prior::T; x; F::Factor
end

function posterior(fx::B200Finite, y::AbstractVector{<:Real})
    F = factor(fx); δ = Float64.(y .- host_mean(fx))
    GC.@preserve δ check(ccall((:sb_factor_set_data, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), ctx().h, F.h, δ))
    B200PosteriorGP(fx.f, fx.x, F)
end

function mean_and_var(fp::B200PosteriorGP, xs)
    ps, vs = components(fp.prior, xs); po, vo = components(fp.prior, fp.x)
    cross, k1 = build_spec(ps, vs, po, vo; symmetric=false)
    # paired (diag) spec: block i with itself
    pd, k2 = build_spec(ps, vs, ps, vs; symmetric=false)   # the library evaluates blocks (i,i) elementwise
    n = sum(length, vs); m = zeros(n); v = zeros(n)
    GC.@preserve k1 k2 check(ccall((:sb_predict, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{SbCovSpec}, Ref{SbCovSpec}, Ptr{Float64}, Ptr{Float64}),
        ctx().h, fp.F.h, cross, pd, m, v))
    m .+ reduce(vcat, [lower(p, x)[1] for (p, x) in zip(ps, vs)]), v
end
mean(fp::B200PosteriorGP, xs) = mean_and_var(fp, xs)[1]
var(fp::B200PosteriorGP, xs) = mean_and_var(fp, xs)[2]

function rand(rng, fx::B200Finite, S::Int)
    F = factor(fx); z = randn(rng, length(fx.x), S); out = similar(z)
    GC.@preserve z out check(ccall((:sb_rand, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}), ctx().h, F.h, z, S, out))
    out .+ host_mean(fx)
end

export b200, B200GPPP

end # module
