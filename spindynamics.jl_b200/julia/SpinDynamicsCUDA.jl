# SpinDynamicsCUDA.jl -- Julia drop-in over libspindyn_cuda (include/spindyn.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: `julia` is not installed in the build image or on
# the GPU box.  The same C ABI is exercised by the Python ctypes mirror (../spindyn), which
# mirrors this file function for function.  Host code stays Julia: random start vectors,
# m x m tridiagonal eigenproblems, Bessel coefficients, kernels and spectrum reconstruction
# are done here exactly as in the reference; every N-length vector operation is a ccall.
#
# Usage (same public API as SpinDynamics.jl):
#     using SpinDynamicsCUDA
#     model = XXZChain(32; Jxy=1.0, Jz=1.0, nup=16)          # GPUModel
#     E0, ψ0 = groundstate(model; lanc_m=30, device=true)
module SpinDynamicsCUDA

using LinearAlgebra, Random
import SpecialFunctions: besselj

export GPUModel, GPUVector, XXZChain, build_model, momenta, nn_hopping, long_range_hopping,
       build_sector_basis, build_full_basis, apply_H!, apply_rescaled_H!, Sz_q_vector,
       lanczos_extremal, lanczos_groundstate, lanczos_groundstate_lean, lanczos_tridiag, estimate_energy_bounds,
       lanczos_sqw, kpm_sqw, krylov_time_evolve, krylov_time_evolve!, KrylovWorkspace, chebyshev_time_evolve,
       ChebyshevWorkspace,
       groundstate, time_evolve, dynamical_structure_factor, neel_state, domain_wall_state, polarized_state,
       polarized_state_with_flips,
       magnetization_per_site, connected_correlations, structure_factor_Sq, structure_factor

const lib = get(ENV, "SPINDYN_CUDA_LIB", "libspindyn_cuda")
const SD_F64, SD_C128 = Cint(0), Cint(1)
const Handle = Ptr{Cvoid}

lasterr() = unsafe_string(ccall((:sd_last_error, lib), Cstring, ()))
function check(rc::Integer)
    rc == 0 && return nothing
    rc == -1 && throw(ArgumentError(lasterr()))
    rc == -3 && throw(OutOfMemoryError())
    error(lasterr())                       # -5 (zero norm) is handled by the callers that can raise it
end

# ------------------------------------------------------------------ context / model
const DEFAULT_CTX = Ref{Handle}(C_NULL)
function default_ctx()
    if DEFAULT_CTX[] == C_NULL
        h = Ref{Handle}()
        check(ccall((:sd_ctx_create, lib), Cint, (Cint, Ref{Handle}), 0, h))
        DEFAULT_CTX[] = h[]
    end
    DEFAULT_CTX[]
end

"""Mirror of SpinModel.Model (SpinModel.jl:6-15) without `states`/`idxmap`."""
mutable struct GPUModel
    h::Handle
    ctx::Handle
    L::Int
    nup::Union{Nothing,Int}
    mode::Symbol
    hopping_list::Vector{Tuple{Int,Int,Float64}}
    onsite_field::Vector{Float64}
    zz_list::Vector{Tuple{Int,Int,Float64}}
    dim::Int
end

function build_model(L::Int; nup::Union{Nothing,Int}=nothing, hopping=Tuple{Int,Int,Float64}[],
                     onsite_field=zeros(L), zz=Tuple{Int,Int,Float64}[], ctx::Handle=default_ctx())
    hop = Vector{Tuple{Int,Int,Float64}}(hopping); zzl = Vector{Tuple{Int,Int,Float64}}(zz)
    fld = Vector{Float64}(onsite_field)
    h = Ref{Handle}()
    check(ccall((:sd_model_create, lib), Cint,
        (Handle, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Float64}, Ref{Handle}),
        ctx, L, nup === nothing ? -1 : nup, hop, length(hop), zzl, length(zzl), fld, h))
    n = Ref{UInt64}()
    check(ccall((:sd_model_dim, lib), Cint, (Handle, Ref{UInt64}), h[], n))
    m = GPUModel(h[], ctx, L, nup, nup === nothing ? :full : :sector, hop, fld, zzl, Int(n[]))
    finalizer(x -> ccall((:sd_model_free, lib), Cint, (Handle,), x.h), m)
end

nn_hopping(L::Int, J::Float64) = [(i, i + 1, J) for i in 1:L-1]
long_range_hopping(L::Int, J::Function) = [(i, j, J(i, j)) for i in 1:L for j in i+1:L]
momenta(m::GPUModel) = [2π * n / m.L for n in 0:m.L-1]

# SpinModel.jl:63-90
function XXZChain(L::Int; Jxy::Real=1.0, Jz::Real=1.0, hz::Real=0.0,
                  nup::Union{Nothing,Int}=nothing, boundary::Symbol=:open)
    hopping = [(i, i + 1, Float64(Jxy) / 2) for i in 1:L-1]
    zz = [(i, i + 1, Float64(Jz)) for i in 1:L-1]
    if boundary == :periodic
        if L > 2
            push!(hopping, (L, 1, Float64(Jxy) / 2)); push!(zz, (L, 1, Float64(Jz)))
        end
    elseif boundary != :open
        throw(ArgumentError("boundary must be :open or :periodic"))
    end
    build_model(L; nup=nup, hopping=hopping, onsite_field=fill(Float64(hz), L), zz=zz)
end

# Basis.jl:23-53: states by device unranking; the Dict is replaced by sd_rank
function states(m::GPUModel, first::Integer=0, count::Integer=m.dim - first)
    out = Vector{UInt64}(undef, count)
    check(ccall((:sd_unrank, lib), Cint, (Handle, UInt64, UInt64, Ptr{UInt64}), m.h, first, count, out))
    out
end
function rank_of(m::GPUModel, s::Vector{UInt64})
    out = Vector{Int64}(undef, length(s))
    check(ccall((:sd_rank, lib), Cint, (Handle, Ptr{UInt64}, UInt64, Ptr{Int64}), m.h, s, length(s), out))
    out
end
build_sector_basis(L::Int, nup::Int) = (m = build_model(L; nup=nup); (states(m), s -> rank_of(m, UInt64[s])[1]))
build_full_basis(L::Int) = (m = build_model(L); (states(m), s -> rank_of(m, UInt64[s])[1]))

# ------------------------------------------------------------------ device vectors
mutable struct GPUVector{T<:Union{Float64,ComplexF64}}
    h::Handle
    m::GPUModel
    owned::Bool
end
dtypeof(::Type{Float64}) = SD_F64
dtypeof(::Type{ComplexF64}) = SD_C128
function GPUVector{T}(m::GPUModel) where T
    h = Ref{Handle}()
    check(ccall((:sd_vec_alloc, lib), Cint, (Handle, Cint, Ref{Handle}), m.h, dtypeof(T), h))
    v = GPUVector{T}(h[], m, true)
    finalizer(x -> x.owned && ccall((:sd_vec_free, lib), Cint, (Handle,), x.h), v)
end
function upload(m::GPUModel, x::Vector{T}) where T<:Union{Float64,ComplexF64}
    length(x) == m.dim || throw(DimensionMismatch("vector does not match the model basis"))
    v = GPUVector{T}(m)
    check(ccall((:sd_vec_upload, lib), Cint, (Handle, Ptr{T}), v.h, x)); v
end
function download(v::GPUVector{T}) where T
    out = Vector{T}(undef, v.m.dim)
    check(ccall((:sd_vec_download, lib), Cint, (Handle, Ptr{T}), v.h, out)); out
end
ondevice(m::GPUModel, x::GPUVector) = x
ondevice(m::GPUModel, x::Vector{<:Real}) = upload(m, Float64.(x))
ondevice(m::GPUModel, x::Vector{<:Complex}) = upload(m, ComplexF64.(x))
function tocomplex(m::GPUModel, x)
    d = ondevice(m, x)
    d isa GPUVector{ComplexF64} && return d
    c = GPUVector{ComplexF64}(m)
    check(ccall((:sd_vec_convert, lib), Cint, (Handle, Handle), c.h, d.h)); c
end
struct SdComplex; re::Float64; im::Float64; end
SdComplex(z::Number) = SdComplex(real(z), imag(z))
scale!(v::GPUVector, s::Number) = (check(ccall((:sd_vec_scale, lib), Cint, (Handle, SdComplex), v.h, SdComplex(s))); v)

# ------------------------------------------------------------------ operator (Hamiltonian.jl)
# apply_H!(out, ψ, model)                                    Hamiltonian.jl:211-273
function apply_H!(out::Vector{T}, ψ::Vector{T}, m::GPUModel) where T<:Union{Float64,ComplexF64}
    @assert length(out) == length(ψ)
    check(ccall((:sd_apply_H_host, lib), Cint, (Handle, Cint, Ptr{T}, Ptr{T}), m.h, dtypeof(T), out, ψ)); out
end
apply_H!(out::GPUVector{T}, ψ::GPUVector{T}, m::GPUModel) where T =
    (check(ccall((:sd_apply_H, lib), Cint, (Handle, Handle, Handle), m.h, out.h, ψ.h)); out)

# apply_rescaled_H!(out, ψ, applyH!, model, a, b)            Hamiltonian.jl:286-301
function apply_rescaled_H!(out::GPUVector{T}, ψ::GPUVector{T}, ::typeof(apply_H!), m::GPUModel,
                           a::Float64, b::Float64) where T
    check(ccall((:sd_apply_rescaled_H, lib), Cint, (Handle, Handle, Handle, Float64, Float64), m.h, out.h, ψ.h, a, b)); out
end

# Sz_q_vector(model, ψ0, q)                                  Hamiltonian.jl:307-337
function Sz_q_vector(m::GPUModel, ψ0, q::Real; device::Bool=false)
    d = ondevice(m, ψ0); ϕ = GPUVector{ComplexF64}(m)
    check(ccall((:sd_szq, lib), Cint, (Handle, Handle, Handle, Float64, Ptr{Float64}), m.h, ϕ.h, d.h, Float64(q), C_NULL))
    device ? ϕ : download(ϕ)
end

# ------------------------------------------------------------------ Lanczos.jl
function lanczos_extremal(::typeof(apply_H!), m::GPUModel; lanc_m::Int=100, tol::Float64=1e-12,
                          rng::AbstractRNG=Random.default_rng(), negate::Bool=false, v0=nothing)
    mm = min(lanc_m, m.dim)
    d0 = v0 === nothing ? upload(m, randn(rng, ComplexF64, m.dim)) : tocomplex(m, v0)
    α = zeros(mm); β = zeros(max(mm - 1, 1)); meff = Ref{Cint}()
    check(ccall((:sd_lanczos_extremal, lib), Cint,
        (Handle, Handle, Cint, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
        m.h, d0.h, lanc_m, tol, negate, α, β, meff))
    k = meff[]
    ev = eigvals(SymTridiagonal(α[1:k], β[1:k-1]))
    minimum(ev), maximum(ev)
end

function estimate_energy_bounds(f::typeof(apply_H!), m::GPUModel; lanc_m::Int=80)   # Lanczos.jl:255-271
    _, Emax = lanczos_extremal(f, m; lanc_m=lanc_m)
    _, Emaxneg = lanczos_extremal(f, m; lanc_m=lanc_m, negate=true)
    -Emaxneg, Emax
end

function lanczos_groundstate(::typeof(apply_H!), m::GPUModel; lanc_m::Int=100, tol::Float64=1e-12,
                             orthogonalize_tol::Float64=1e-10, rng::AbstractRNG=Random.default_rng(),
                             v0=nothing, device::Bool=false)
    mm = min(lanc_m, m.dim)
    d0 = v0 === nothing ? upload(m, randn(rng, Float64, m.dim)) : ondevice(m, v0)
    α = zeros(mm); β = zeros(max(mm - 1, 1)); mact = Ref{Cint}(); V = Ref{Handle}()
    check(ccall((:sd_lanczos_groundstate, lib), Cint,
        (Handle, Handle, Cint, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Cint}, Ref{Handle}),
        m.h, d0.h, lanc_m, tol, orthogonalize_tol, α, β, mact, V))
    ma = mact[]
    F = eigen(SymTridiagonal(α[1:ma], β[1:min(ma - 1, mm - 1)]))            # Lanczos.jl:164-165
    i = argmin(F.values); y = ComplexF64.(F.vectors[:, i])
    ψ = GPUVector{Float64}(m); n2 = Ref{Float64}()
    check(ccall((:sd_lincomb, lib), Cint, (Handle, Ptr{ComplexF64}, Cint, Handle, Ref{Float64}), V[], y, ma, ψ.h, n2))
    scale!(ψ, 1 / sqrt(n2[]))
    ccall((:sd_vecset_free, lib), Cint, (Handle,), V[])
    F.values[i], (device ? ψ : download(ψ))
end

# Extension (SURVEY 8f-3, not in the reference): ground state on three device vectors instead of the N x m basis --
# the three-term recurrence twice from the same v0, pass 2 accumulating the Ritz vector (sd_lanczos_lean).
function lanczos_groundstate_lean(::typeof(apply_H!), m::GPUModel; lanc_m::Int=100, tol::Float64=1e-12,
                                  rng::AbstractRNG=Random.default_rng(), v0=nothing, device::Bool=false)
    mm = min(lanc_m, m.dim)
    d0 = v0 === nothing ? upload(m, randn(rng, Float64, m.dim)) : ondevice(m, v0)
    α = zeros(mm); β = zeros(max(mm, 1)); meff = Ref{Cint}()
    check(ccall((:sd_lanczos_lean, lib), Cint,
        (Handle, Handle, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Cint}, Ptr{Float64}, Handle, Ptr{Float64}),
        m.h, d0.h, mm, tol, α, β, meff, C_NULL, C_NULL, C_NULL))
    k = meff[]
    F = eigen(SymTridiagonal(α[1:k], β[1:k-1])); i = argmin(F.values); y = Float64.(F.vectors[:, i])
    ψ = GPUVector{Float64}(m); n2 = Ref{Float64}()
    check(ccall((:sd_lanczos_lean, lib), Cint,
        (Handle, Handle, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Cint}, Ptr{Float64}, Handle, Ref{Float64}),
        m.h, d0.h, k, tol, α, β, meff, y, ψ.h, n2))
    scale!(ψ, 1 / sqrt(n2[]))
    F.values[i], (device ? ψ : download(ψ))
end

function lanczos_tridiag(::typeof(apply_H!), m::GPUModel, v; lanc_m::Int=100, tol::Float64=1e-12)
    dv = tocomplex(m, v); mm = min(lanc_m, m.dim)
    α = zeros(mm); β = zeros(max(mm - 1, 1)); meff = Ref{Cint}(); nv = Ref{Float64}()
    rc = ccall((:sd_lanczos_tridiag, lib), Cint,
        (Handle, Handle, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Cint}, Ref{Float64}),
        m.h, dv.h, lanc_m, tol, α, β, meff, nv)
    rc == -5 && error("starting vector has zero norm")                         # Lanczos.jl:210-212
    check(rc)
    α[1:meff[]], β[1:meff[]-1], nv[]
end

# ------------------------------------------------------------------ LanczosSqw.jl / KPM_Sqw.jl
function spectral_from_tridiagonal(α, β, norm_phi, E0, ω; eta=0.05, broaden=:lorentz)
    F = eigen(SymTridiagonal(α, β)); w = abs2.(F.vectors[1, :]) .* norm_phi^2
    shifted = ω .- (F.values .- E0)'
    broaden == :lorentz ? vec(((1 / pi) .* (eta ./ (shifted .^ 2 .+ eta^2))) * w) :
    broaden == :gauss ? vec(((1 / (sqrt(2pi) * eta)) .* exp.(-(shifted .^ 2) ./ (2eta^2))) * w) :
    error("unknown broadening: $broaden")
end

# The multi-vector kernel wins where one vector cannot fill the GPU (L = 16) and loses to the block kernel once that is
# bandwidth-bound (L = 28): the automatic choice batches small bases only.
const Q_BATCH_AUTO_MAX_DIM = 1 << 20
# q_batch: all momenta as one interleaved [state][q] multi-vector (sd_lanczos_tridiag_szq_batch, at most 128 per call):
# the reference's Threads.@threads q-loop as data parallelism, two fused kernels per Lanczos step for ALL momenta.
function lanczos_sqw(ψ0, m::GPUModel, q_list::AbstractVector{Float64}, ω::AbstractVector{Float64};
                     lanc_m::Int=200, eta::Float64=0.05, broaden::Symbol=:lorentz, q_batch::Bool=length(q_list) >= 2 && m.dim <= Q_BATCH_AUTO_MAX_DIM)
    ψc = tocomplex(m, ψ0); tmp = GPUVector{ComplexF64}(m); apply_H!(tmp, ψc, m)
    r = Ref{SdComplex}(); check(ccall((:sd_vec_dotu, lib), Cint, (Handle, Handle, Ref{SdComplex}), ψc.h, tmp.h, r))
    E0 = r[].re                                                               # LanczosSqw.jl:59
    S = zeros(length(q_list), length(ω))
    if q_batch
        for lo in 1:128:length(q_list)
            qs = collect(q_list[lo:min(lo + 127, end)]); n = length(qs)
            α = zeros(lanc_m, n); β = zeros(lanc_m, n); meff = zeros(Cint, n); nϕ = zeros(n)    # column c = momentum c
            check(ccall((:sd_lanczos_tridiag_szq_batch, lib), Cint,
                (Handle, Handle, Ptr{Float64}, Cint, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}, Ptr{Float64}),
                m.h, ψc.h, qs, n, lanc_m, 1e-12, α, β, meff, nϕ))
            for c in 1:n
                k = meff[c]; k == 0 && continue                               # LanczosSqw.jl:69-72
                S[lo + c - 1, :] .= spectral_from_tridiagonal(α[1:k, c], β[1:k-1, c], nϕ[c], E0, ω; eta=eta, broaden=broaden)
            end
        end
        return S
    end
    ϕ = GPUVector{ComplexF64}(m)
    for (iq, q) in enumerate(q_list)                                          # sequential device work
        n2 = Ref{Float64}()
        check(ccall((:sd_szq, lib), Cint, (Handle, Handle, Handle, Float64, Ref{Float64}), m.h, ϕ.h, ψc.h, q, n2))
        n2[] == 0 && continue
        α, β, nϕ = lanczos_tridiag(apply_H!, m, ϕ; lanc_m=lanc_m)
        S[iq, :] .= spectral_from_tridiagonal(α, β, nϕ, E0, ω; eta=eta, broaden=broaden)
    end
    S
end

rescaling_from_bounds(Emin, Emax) = ((Emax - Emin) / (2 * 0.99), (Emax + Emin) / 2)   # KPM_Sqw.jl:13-17
function jackson(M)                                                                    # KPM_Sqw.jl:131-137
    [((M - n + 1) * cos(pi * n / (M + 1)) + sin(pi * n / (M + 1)) * cot(pi / (M + 1))) / (M + 1) for n in 0:M-1]
end
function compute_chebyshev_moments(::typeof(apply_H!), ϕ::GPUVector{ComplexF64}, M::Int, a, b, m::GPUModel)
    μ = zeros(M)
    check(ccall((:sd_kpm_moments, lib), Cint, (Handle, Handle, Cint, Float64, Float64, Ptr{Float64}), m.h, ϕ.h, M, a, b, μ)); μ
end
function kpm_sqw(ψ0, m::GPUModel, q_list::AbstractVector{Float64}, ω::AbstractVector{Float64};
                 a=nothing, b=nothing, kpm_m::Int=200, kernel::Symbol=:jackson, q_batch::Bool=length(q_list) >= 2 && m.dim <= Q_BATCH_AUTO_MAX_DIM)
    ψc = tocomplex(m, ψ0); tmp = GPUVector{ComplexF64}(m); r = Ref{SdComplex}()
    check(ccall((:sd_apply_H_dot, lib), Cint, (Handle, Handle, Handle, Ref{SdComplex}), m.h, tmp.h, ψc.h, r))
    E0 = r[].re
    if a === nothing || b === nothing
        a, b = rescaling_from_bounds(estimate_energy_bounds(apply_H!, m)...)
    end
    g = kernel == :jackson ? jackson(kpm_m) : kernel == :lorentz ? [sinh(3.0 * (1 - n / kpm_m)) / sinh(3.0) for n in 0:kpm_m-1] : ones(kpm_m)
    S = zeros(length(q_list), length(ω))
    function reconstruct!(iq, μ, nϕ)                                          # KPM_Sqw.jl:58-90
        for (iw, w) in enumerate(ω)
            x = (w + E0 - b) / a
            abs(x) >= 1 && continue
            Tm2, Tm1 = 1.0, x; s = μ[1] + (kpm_m >= 2 ? 2μ[2] * x : 0.0)
            for n in 3:kpm_m
                Tn = 2x * Tm1 - Tm2; s += 2μ[n] * Tn; Tm2, Tm1 = Tm1, Tn
            end
            S[iq, iw] = nϕ^2 * max(0.0, s / (a * pi * sqrt(1 - x^2)))
        end
    end
    batched = q_batch
    if q_batch                                                                # sd_kpm_moments_szq_batch, at most 128 momenta per call
        for lo in 1:128:length(q_list)
            qs = collect(q_list[lo:min(lo + 127, end)]); n = length(qs)
            μ = zeros(kpm_m, n); nϕ = zeros(n); blown = Ref{Cint}(0)
            check(ccall((:sd_kpm_moments_szq_batch, lib), Cint,
                (Handle, Handle, Ptr{Float64}, Cint, Cint, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
                m.h, ψc.h, qs, n, kpm_m, a, b, μ, nϕ, blown))
            if blown[] != 0                                                   # KPM_Sqw.jl:117-121 applies: per-momentum path
                batched = false; fill!(S, 0.0); break
            end
            for c in 1:n
                nϕ[c] == 0 && continue
                reconstruct!(lo + c - 1, μ[:, c] .* g, nϕ[c])
            end
        end
    end
    batched && return S
    ϕ = GPUVector{ComplexF64}(m)
    for (iq, q) in enumerate(q_list)
        n2 = Ref{Float64}()
        check(ccall((:sd_szq, lib), Cint, (Handle, Handle, Handle, Float64, Ref{Float64}), m.h, ϕ.h, ψc.h, q, n2))
        nϕ = sqrt(n2[]); nϕ == 0 && continue
        scale!(ϕ, 1 / nϕ)
        reconstruct!(iq, compute_chebyshev_moments(apply_H!, ϕ, kpm_m, a, b, m) .* g, nϕ)
    end
    S
end

# ------------------------------------------------------------------ TimeEvolution
function krylov_time_evolve(ψ0, dt::Float64, ::typeof(apply_H!), m::GPUModel; kry_m::Int=30, device::Bool=false)
    d0 = ondevice(m, ψ0)
    α = zeros(ComplexF64, kry_m); β = zeros(max(kry_m - 1, 1)); meff = Ref{Cint}(); n0 = Ref{Float64}(); V = Ref{Handle}()
    check(ccall((:sd_krylov_basis, lib), Cint,
        (Handle, Handle, Cint, Ptr{ComplexF64}, Ptr{Float64}, Ref{Cint}, Ref{Float64}, Ref{Handle}),
        m.h, d0.h, kry_m, α, β, meff, n0, V))
    n0[] == 0 && return ψ0
    k = meff[]
    F = eigen(Matrix(Tridiagonal(ComplexF64.(β[1:k-1]), α[1:k], ComplexF64.(β[1:k-1]))))   # Krylov.jl:175-176
    y = F.vectors * Diagonal(exp.(-1im .* F.values .* dt)) * F.vectors' * [n0[]; zeros(k - 1)]
    ψt = GPUVector{ComplexF64}(m); n2 = Ref{Float64}()
    check(ccall((:sd_lincomb, lib), Cint, (Handle, Ptr{ComplexF64}, Cint, Handle, Ref{Float64}), V[], ComplexF64.(y), k, ψt.h, n2))
    scale!(ψt, 1 / sqrt(n2[]))
    ccall((:sd_vecset_free, lib), Cint, (Handle,), V[])
    device ? ψt : download(ψt)
end

# Krylov.jl:25-118: the in-place twin.  The Krylov basis is device-resident inside the library for one call, so the
# workspace only carries the sizes the reference asserts on.
struct KrylovWorkspace; n::Int; m::Int; end
function krylov_time_evolve!(ψ_out::Vector{ComplexF64}, ψ_in::Vector{ComplexF64}, dt::Float64, f::typeof(apply_H!),
                             m::GPUModel, ws::KrylovWorkspace; kry_m::Int=30)
    @assert length(ψ_out) == length(ψ_in)
    @assert ws.m ≥ kry_m
    if norm(ψ_in) == 0
        copyto!(ψ_out, ψ_in); return ψ_out
    end
    copyto!(ψ_out, krylov_time_evolve(ψ_in, dt, f, m; kry_m=kry_m))
    return nothing
end

# Chebyshev.jl:19-36: ϕ_prev/ϕ_curr/ϕ_next/ψ_t live on the device inside sd_chebyshev_evolve; N is kept for the size check
struct ChebyshevWorkspace{T<:Number}; N::Int; end
ChebyshevWorkspace{T}(ψ::AbstractVector) where T = ChebyshevWorkspace{T}(length(ψ))
ChebyshevWorkspace(ψ::AbstractVector{T}) where T<:Number = ChebyshevWorkspace{T}(length(ψ))

function chebyshev_time_evolve(ψ0, dt::Float64, ::typeof(apply_H!), m::GPUModel; cheb_n::Int=100,
                               Ebounds::Tuple{Float64,Float64}=(-1.0, 1.0), device::Bool=false, workspace=nothing)
    @assert cheb_n >= 1 "cheb_n must be >= 1"
    workspace === nothing || @assert workspace.N == length(ψ0) "Workspace size mismatch"
    Emin, Emax = Ebounds; a = (Emax - Emin) / (2 * 0.9999); b = (Emax + Emin) / 2           # Chebyshev.jl:70-79
    c = [(2 - (k == 0)) * (-1im)^k * besselj(k, a * dt) * exp(-1im * b * dt) for k in 0:cheb_n-1]
    d0 = ondevice(m, ψ0); d0 isa GPUVector{ComplexF64} || throw(InexactError(:chebyshev_time_evolve, ComplexF64, 0))
    out = GPUVector{ComplexF64}(m)
    check(ccall((:sd_chebyshev_evolve, lib), Cint, (Handle, Handle, Ptr{ComplexF64}, Cint, Float64, Float64, Handle),
                m.h, d0.h, ComplexF64.(c), cheb_n, a, b, out.h))
    device ? out : download(out)
end

# ------------------------------------------------------------------ PublicAPI.jl
groundstate(m::GPUModel; method::Symbol=:lanczos, kw...) =
    method === :lanczos ? lanczos_groundstate(apply_H!, m; kw...) : throw(ArgumentError("unsupported ground-state method: $method"))
function time_evolve(m::GPUModel, ψ0, t::Real; method::Symbol=:krylov, Ebounds=nothing, kw...)
    method === :krylov && return krylov_time_evolve(ψ0, Float64(t), apply_H!, m; kw...)
    method === :chebyshev && return chebyshev_time_evolve(ψ0, Float64(t), apply_H!, m;
        Ebounds=Ebounds === nothing ? estimate_energy_bounds(apply_H!, m) : Ebounds, kw...)
    throw(ArgumentError("unsupported time-evolution method: $method"))
end
function dynamical_structure_factor(m::GPUModel, ψ0, q, ω; method::Symbol=:lanczos, kw...)
    method === :lanczos && return lanczos_sqw(ψ0, m, Float64.(q), Float64.(ω); kw...)
    method === :kpm && return kpm_sqw(ψ0, m, Float64.(q), Float64.(ω); kw...)
    throw(ArgumentError("unsupported dynamical structure-factor method: $method"))
end

# ------------------------------------------------------------------ Observables.jl:14-109 (ψ may stay on the device)
function observables(ψ, m::GPUModel)
    d = ondevice(m, ψ); mags = zeros(m.L); zz = zeros(m.L)
    check(ccall((:sd_vec_observables, lib), Cint, (Handle, Ptr{Float64}, Ptr{Float64}), d.h, mags, zz))
    mags, zz
end
magnetization_per_site(ψ, m::GPUModel) = observables(ψ, m)[1]
function connected_correlations(ψ, m::GPUModel)
    mags, zz = observables(ψ, m); L = m.L
    [(zz[r+1] - sum(mags[i] * mags[mod1(i + r, L)] for i in 1:L)) / L for r in 0:L-1]
end
function structure_factor_Sq(ψ, m::GPUModel)                       # Observables.jl:101-109 (naive DFT: L numbers)
    C = connected_correlations(ψ, m); L = m.L
    Dict(2π * (n - 1) / L => real(sum(C[r+1] * exp(-2π * im * (n - 1) * r / L) for r in 0:L-1)) for n in 1:L)
end
structure_factor(m::GPUModel, ψ) = structure_factor_Sq(ψ, m)       # PublicAPI.jl:94-106

# InitialStates.jl through sd_rank (get(idxmap, s, 0) / Int(s)+1 without a Dict); host Vector{Float64} like the reference
function onehot(m::GPUModel, s::UInt64, what::String)
    idx = rank_of(m, UInt64[s])[1]
    idx == 0 && throw(ArgumentError("$what is not contained in the model basis"))
    ψ = zeros(m.dim); ψ[idx] = 1.0; ψ
end
function domain_wall_state(m::GPUModel)                             # InitialStates.jl:9-34
    nup = m.nup === nothing ? Int(ceil(m.L / 2)) : m.nup
    s = UInt64(0); for i in 0:nup-1; s |= UInt64(1) << i; end
    onehot(m, s, "domain-wall state")
end
function neel_state(m::GPUModel)                                    # :40-63
    s = UInt64(0); for i in 1:m.L; isodd(i) && (s |= UInt64(1) << (i - 1)); end
    onehot(m, s, "Neel state")
end
polarized_state(m::GPUModel; up::Bool=true) =                       # :70-90
    onehot(m, up ? (UInt64(1) << m.L) - UInt64(1) : UInt64(0), "requested polarized state")
function polarized_state_with_flips(m::GPUModel, flips::Vector{Int})   # :97-130
    for site in flips
        1 <= site <= m.L || throw(ArgumentError("flip site $site is outside the model with L=$(m.L)"))
    end
    s = (UInt64(1) << m.L) - UInt64(1)
    for site in flips; s ⊻= UInt64(1) << (site - 1); end
    onehot(m, s, "requested flipped polarized state")
end

end # module
