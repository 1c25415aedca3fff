# ToyFHEB200.jl -- the reference-side binding of libtoyfhe_b200.so.
#
# UNTESTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  The
# same entry points are exercised through ctypes by tests/ (Python), which mirror
# these methods one to one.  `include` this file after `using ToyFHE`; it adds
# more specific methods at the dispatch points the engine replaces:
#
#   NTT.nntt / NTT.inntt for StructArray-of-CRTEncoded storage   (src/crt.jl:247-267)
#   ToyFHE.modswitch(::RingElement)                               (src/crt.jl:226-228)
#   ToyFHE.enc_mul for RNS ciphertexts                            (src/rlwe_she.jl:247-262)
#
# Residues: GaloisFields.PrimeField{I,p} is an isbits wrapper of one integer `n`
# in [0,p); for p < 2^62 the engine needs I == Int64/UInt64 so that each field
# array of the StructArray is a contiguous N x 8-byte buffer.
module ToyFHEB200

using ToyFHE, StructArrays, OffsetArrays
using ToyFHE: CRTEncoded
using ToyFHE.NTT: NegacyclicRing, RingCoeffs, RingElement, degree, coeffs_primal
import ToyFHE.NTT: nntt, inntt

const LIB = get(ENV, "TOYFHE_B200_LIB", joinpath(@__DIR__, "..", "toyfhe.jl_b200", "lib", "libtoyfhe_b200.so"))

struct EngineError <: Exception
    code::Cint
    msg::String
end
function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:tfb_last_error, LIB), Cstring, ()))
    rc == 1 ? throw(ArgumentError(msg)) : throw(EngineError(rc, msg))   # TFB_EINVAL <-> UsageError/@assert
end

# ---- one engine context per ring instance (the ring is a type parameter in ToyFHE) ----
const CONTEXTS = IdDict{Any,Ptr{Cvoid}}()
moduli_of(::Type{<:CRTEncoded{L,M}}) where {L,M} = UInt64[UInt64(ToyFHE.modulus(T)) for T in M.parameters]
function context(ℛ::NegacyclicRing{F}) where {F<:CRTEncoded}
    get!(CONTEXTS, ℛ) do
        q = moduli_of(F)
        ψ = UInt64[UInt64(c.n) for c in ℛ.ψ.c]            # per-prime minimal primitive 2N-th roots (crt.jl:293)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tfb_ctx_create, LIB), Cint, (Cint, UInt32, UInt32, Ptr{UInt64}, Ptr{UInt64}, Ref{Ptr{Cvoid}}),
                    0, degree(ℛ), length(q), q, ψ, out))
        out[]
    end
end

# pack the L field arrays into one [L][N] buffer (residue-major, the engine's layout) and back
function pack(sa::StructArray)
    fa = fieldarrays(sa); N = length(sa)
    buf = Matrix{UInt64}(undef, N, length(fa))
    for (i, a) in enumerate(fa)
        copyto!(view(buf, :, i), reinterpret(UInt64, a))
    end
    buf
end
function unpack(sa::StructArray, buf::Matrix{UInt64})
    StructArray{eltype(sa)}(tuple((collect(reinterpret(eltype(a), buf[:, i])) for (i, a) in enumerate(fieldarrays(sa)))...))
end

for (jl, sym) in ((:nntt, :tfb_ntt_fwd_host), (:inntt, :tfb_ntt_inv_host))
    @eval function $jl(rcs::RingCoeffs{ℛ,T,OffsetVector{T,S}})::RingCoeffs{ℛ} where {ℛ,T<:CRTEncoded,S<:StructArray{T}}
        oa = rcs.coeffs
        buf = pack(oa.parent)
        check(ccall(($(QuoteNode(sym)), LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                    context(ℛ), buf, buf, size(buf, 2), C_NULL))
        RingCoeffs{ℛ}(OffsetArray(unpack(oa.parent, buf), axes(oa)...))
    end
end

# CKKS rescale / special-prime contract: crt.jl:215-228
function ToyFHE.modswitch(re::RingElement{ℛ,Field}) where {ℛ,Field<:CRTEncoded}
    p = coeffs_primal(re)
    inbuf = pack(p.parent)
    outbuf = Matrix{UInt64}(undef, size(inbuf, 1), size(inbuf, 2) - 1)
    check(ccall((:tfb_rescale_host, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), inbuf, outbuf, 1, C_NULL))
    ℛ′ = ToyFHE.drop_last(ℛ)
    T′ = eltype(ℛ′)
    fa = fieldarrays(p.parent)[1:end-1]
    sa = StructArray{T′}(tuple((collect(reinterpret(eltype(a), outbuf[:, i])) for (i, a) in enumerate(fa))...))
    RingElement{ℛ′}(OffsetArray(sa, axes(p)...), nothing)
end

# Whole-ciphertext products: one call per product instead of 7 forward + 4 inverse transforms.
# c1, c2: 2-component ciphertexts over the same RNS ring (CKKS/BGV form: no basis change).
function ct_tensor(ℛ, c1::Vector{<:RingElement}, c2::Vector{<:RingElement})
    @assert length(c1) == 2 && length(c2) == 2
    a = cat((pack(coeffs_primal(x).parent) for x in c1)...; dims=3)   # [N, L, 2]
    b = cat((pack(coeffs_primal(x).parent) for x in c2)...; dims=3)
    out = Array{UInt64}(undef, size(a, 1), size(a, 2), 3)
    check(ccall((:tfb_ct_tensor_host, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), a, b, out, 1, C_NULL))
    proto = coeffs_primal(c1[1])
    [RingElement{ℛ}(OffsetArray(unpack(proto.parent, out[:, :, k]), axes(proto)...), nothing) for k in 1:3]
end

# BFV: expand to ℛbig, tensor, scale-and-round, contract (bfv.jl:34-40 hooks folded into one call)
function bfv_mul(ℛ, ℛbig, t::Integer, c1::Vector{<:RingElement}, c2::Vector{<:RingElement})
    a = cat((pack(coeffs_primal(x).parent) for x in c1)...; dims=3)
    b = cat((pack(coeffs_primal(x).parent) for x in c2)...; dims=3)
    out = Array{UInt64}(undef, size(a, 1), size(a, 2), 3)
    check(ccall((:tfb_bfv_mul_host, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), context(ℛbig), UInt64(t), a, b, out, 1, C_NULL))
    proto = coeffs_primal(c1[1])
    [RingElement{ℛ}(OffsetArray(unpack(proto.parent, out[:, :, k]), axes(proto)...), nothing) for k in 1:3]
end

# BFV plaintext maps (bfv.jl:21-29): Delta * m and mod(divround(SignedMod(x), Delta), t), exact on the device
limbs(x::Integer) = (n = cld(max(ndigits(x, base=2), 1), 64); UInt64[UInt64((x >> (64 * (i - 1))) & typemax(UInt64)) for i in 1:n])

function bfv_encode(ℛ, t::Integer, Δ::Integer, m::Vector{UInt64})
    d = limbs(Δ)
    L = length(ℛ.ψ.c)                                   # number of RNS primes (crt.jl:293)
    out = Matrix{UInt64}(undef, length(m), L)
    check(ccall((:tfb_bfv_encode_host, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ptr{UInt64}, UInt32, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), UInt64(t), d, length(d), m, out, 1, C_NULL))
    out
end

function bfv_decode(ℛ, t::Integer, Δ::Integer, b::RingElement)
    d = limbs(Δ)
    buf = pack(coeffs_primal(b).parent)
    out = Vector{UInt64}(undef, size(buf, 1))
    check(ccall((:tfb_bfv_decode_host, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ptr{UInt64}, UInt32, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), UInt64(t), d, length(d), buf, out, 1, C_NULL))
    out
end

end # module
