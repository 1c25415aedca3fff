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
using ToyFHE: NTT
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

# ---- keyswitch with a device-resident evaluation key (rlwe_she.jl:315-347; INTEGRATION.md section 3) ----
# The evaluation key is uploaded ONCE in the NTT domain ([D][2][L'][N], component 1 = mask, 2 = masked) -- what the
# reference caches in key.mask.dual / key.masked.dual after first use (pow2_cyc_rings.jl:132-138) -- and every
# keyswitch then moves only the ciphertext (2-3 polynomials in, 2 out).
struct DeviceKey
    ctx::Ptr{Cvoid}          # ring of the ciphertext
    ext::Ptr{Cvoid}          # raised ring (ModulusRaised) or C_NULL
    ptr::Ptr{Cvoid}          # device buffer [D][2][L'][N]
    D::Int
    w::Int                   # relin_window (0 = CRT digits)
end

function dmalloc(ctx, bytes)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tfb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx, bytes, p))
    p[]
end
h2d(ctx, dst, src::Array{UInt64}) = check(ccall((:tfb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Csize_t, Ptr{Cvoid}), ctx, dst, src, sizeof(src), C_NULL))
d2h(ctx, dst::Array{UInt64}, src) = check(ccall((:tfb_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), ctx, dst, src, sizeof(dst), C_NULL))

# ek.key :: Vector of (mask, masked) RingElements over ℛkey (rlwe_she.jl:273-298); ℛ = ciphertext ring
function DeviceKey(ℛ, ℛkey, ek, w::Integer; raised::Bool = ℛkey !== ℛ)
    kctx = context(ℛkey)
    D = length(ek.key)
    host = cat((cat(pack(NTT.coeffs_dual(k.mask).parent), pack(NTT.coeffs_dual(k.masked).parent); dims=3) for k in ek.key)...; dims=4)  # [N, L', 2, D]
    dev = dmalloc(kctx, sizeof(host))
    h2d(kctx, dev, host)
    DeviceKey(context(ℛ), raised ? kctx : C_NULL, dev, D, w)
end

# c :: Vector of 2 or 3 RingElements over ℛ; returns the 2 components of keyswitch(ek, c)
function keyswitch(ℛ, dk::DeviceKey, c::Vector{<:RingElement})
    comps = length(c)
    host = cat((pack(coeffs_primal(x).parent) for x in c)...; dims=3)       # [N, L, comps]
    N, L = size(host, 1), size(host, 2)
    din, dout = dmalloc(dk.ctx, sizeof(host)), dmalloc(dk.ctx, N * L * 2 * 8)
    h2d(dk.ctx, din, host)
    check(ccall((:tfb_keyswitch, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt64, Ptr{Cvoid}),
                dk.ctx, dk.ext, dk.w, dk.ptr, dk.D, din, comps, dout, 1, C_NULL))
    out = Array{UInt64}(undef, N, L, 2)
    d2h(dk.ctx, out, dout)
    check(ccall((:tfb_sync, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dk.ctx, C_NULL))
    for p in (din, dout)
        check(ccall((:tfb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dk.ctx, p))
    end
    proto = coeffs_primal(c[1])
    [RingElement{ℛ}(OffsetArray(unpack(proto.parent, out[:, :, k]), axes(proto)...), nothing) for k in 1:2]
end

# rotate(gk, c) = keyswitch(gk, apply_galois_element(c, g)) (rlwe_she.jl:355-359): the automorphism is
# tfb_galois on the device buffer between the upload and tfb_keyswitch; same data movement as above.

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
