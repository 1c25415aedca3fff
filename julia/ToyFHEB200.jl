# ToyFHEB200.jl -- the reference-side binding of libtoyfhe_b200.so.
#
# NOT RUN IN THIS REPOSITORY: the build image has no Julia toolchain.  What IS checked here:
#   * tests/test_julia_shim.py parses every `ccall` below and compares symbol, argument count and argument types with the
#     prototypes in include/toyfhe_b200.h (CPU test), and
#   * tests/test_gpu_julia_mirror.py replays each override's call sequence through raw ctypes -- same symbols, same
#     argument order, same column-major buffer shapes -- and checks the results against the oracle (GPU test).
#
# `include` this file after `using ToyFHE`.  It adds MORE SPECIFIC METHODS at the dispatch points the engine replaces, so
# the scheme layer (keygen / encrypt / decrypt / * / + / keyswitch / rotate / modswitch) is untouched:
#
#   NTT.nntt / NTT.inntt for StructArray-of-CRTEncoded storage         src/crt.jl:247-267
#   ToyFHE.modswitch(::RingElement)            (CKKS rescale)           src/crt.jl:226-228
#   ToyFHE.enc_mul(c1, c2) for 2-component RNS ciphertexts              src/rlwe_she.jl:247-262  (`c1*c2`, :264-266)
#        BFVParams  -> tfb_bfv_mul_host   (mul_expand / mul_contract hooks of src/bfv.jl:34-40 folded in)
#        otherwise  -> tfb_ct_tensor_host (CKKS / BGV: default hooks)
#   ToyFHE.keyswitch(::KeySwitchKey, ::CipherText) for RNS ciphertexts  src/rlwe_she.jl:315-349, src/modulusraising.jl:35-49
#        with the evaluation key uploaded once per (key, level) in the NTT domain
#   NTT.apply_galois_element(::RingElement, g)                          src/pow2_cyc_rings.jl:321-329
#
# Residues: GaloisFields.PrimeField{I,p} is an isbits wrapper of one integer `n` in [0,p); for p < 2^62 the engine needs
# I == Int64/UInt64 so that each field array of the StructArray is a contiguous N x 8-byte buffer.
module ToyFHEB200

using ToyFHE, StructArrays, OffsetArrays
using ToyFHE: CRTEncoded, CipherText, KeySwitchKey, KeyComponent, BFVParams, ModulusRaised, SHEShemeParams
using ToyFHE: NTT
using ToyFHE.NTT: NegacyclicRing, RingCoeffs, RingElement, degree, coeffs_primal, coeffs_dual
import ToyFHE.NTT: nntt, inntt

const LIB = get(ENV, "TOYFHE_B200_LIB", joinpath(@__DIR__, "..", "toyfhe.jl_b200", "lib", "libtoyfhe_b200.so"))
const DEVICE = parse(Cint, get(ENV, "TOYFHE_B200_DEVICE", "0"))      # one process per GPU: the rank's device

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
                    DEVICE, degree(ℛ), length(q), q, ψ, out))
        out[]
    end
end
nprimes(::NegacyclicRing{F}) where {L,M,F<:CRTEncoded{L,M}} = L

# ---- host staging: one reusable [N, L, k] buffer per shape instead of a fresh Matrix per call ----
const STAGING = Dict{Tuple{Int,Int,Int,Symbol},Array{UInt64,3}}()
staging(N, L, k, tag::Symbol) = get!(() -> Array{UInt64,3}(undef, N, L, k), STAGING, (N, L, k, tag))

# residue-major packing of the L field arrays of a StructArray into column j of buf[:, :, j] and back
function pack!(buf::Array{UInt64,3}, j::Int, sa::StructArray)
    for (i, a) in enumerate(fieldarrays(sa))
        copyto!(view(buf, :, i, j), reinterpret(UInt64, a))
    end
    buf
end
function unpack(proto::StructArray, buf::AbstractMatrix{UInt64}, ::Type{T} = eltype(proto)) where {T}
    fa = fieldarrays(proto)
    StructArray{T}(tuple((collect(reinterpret(eltype(fa[i]), buf[:, i])) for i in 1:size(buf, 2))...))
end
element(::Type{RingElement{ℛ}}, proto, buf) where {ℛ} = RingElement{ℛ}(OffsetArray(unpack(proto.parent, buf), axes(proto)...), nothing)

# ---- transforms: crt.jl:247-267 ----
for (jl, sym) in ((:nntt, :tfb_ntt_fwd_host), (:inntt, :tfb_ntt_inv_host))
    @eval function $jl(rcs::RingCoeffs{ℛ,T,OffsetVector{T,S}})::RingCoeffs{ℛ} where {ℛ,T<:CRTEncoded,S<:StructArray{T}}
        oa = rcs.coeffs
        N, L = length(oa), nprimes(ℛ)
        buf = pack!(staging(N, L, 1, :ntt), 1, oa.parent)
        check(ccall(($(QuoteNode(sym)), LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                    context(ℛ), buf, buf, L, C_NULL))
        RingCoeffs{ℛ}(OffsetArray(unpack(oa.parent, view(buf, :, :, 1)), axes(oa)...))
    end
end

# ---- CKKS rescale / special-prime contract: crt.jl:215-228 ----
function ToyFHE.modswitch(re::RingElement{ℛ,Field}) where {ℛ,Field<:CRTEncoded}
    p = coeffs_primal(re)
    N, L = length(p), nprimes(ℛ)
    inbuf = pack!(staging(N, L, 1, :rs_in), 1, p.parent)
    outbuf = staging(N, L - 1, 1, :rs_out)
    check(ccall((:tfb_rescale_host, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), inbuf, outbuf, 1, C_NULL))
    ℛ′ = ToyFHE.drop_last(ℛ)
    fa = fieldarrays(p.parent)[1:end-1]
    sa = StructArray{eltype(ℛ′)}(tuple((collect(reinterpret(eltype(a), outbuf[:, i, 1])) for (i, a) in enumerate(fa))...))
    RingElement{ℛ′}(OffsetArray(sa, axes(p)...), nothing)
end

# ---- ciphertext * ciphertext: rlwe_she.jl:247-266 ----
# `c1 * c2` calls enc_mul(c1, c2); this method is more specific than the reference's untyped enc_mul(c1, c2) for two
# 2-component ciphertexts over the same RNS ring, i.e. every product the reference's tests and examples form before
# relinearisation.  Longer ciphertexts fall through to the reference method (whose ring products still reach the engine
# through nntt / inntt above).
const RNSCipher{P,ℛ} = CipherText{<:Any,P,<:RingElement{ℛ,<:CRTEncoded},2}

plain_modulus(params::BFVParams) = ToyFHE.modulus(ToyFHE.NTT.base_ring(params.ℛplain))       # bfv.jl:38

function ToyFHE.enc_mul(c1::RNSCipher{P,ℛ}, c2::RNSCipher{P,ℛ}) where {P<:SHEShemeParams,ℛ}
    c1.params !== c2.params && throw(ToyFHE.UsageError("Attempting to multiply ciphertexts with differing parameters"))
    params = c1.params
    proto = coeffs_primal(c1[1])
    N, L = length(proto), nprimes(ℛ)
    a, b, out = staging(N, L, 2, :mul_a), staging(N, L, 2, :mul_b), staging(N, L, 3, :mul_out)
    for j in 1:2
        pack!(a, j, coeffs_primal(c1[j]).parent)
        pack!(b, j, coeffs_primal(c2[j]).parent)
    end
    if params isa BFVParams
        # mul_expand (switch to ℛbig, bfv.jl:34), tensor, mul_contract (multround + switch back, bfv.jl:35-40) in one call
        check(ccall((:tfb_bfv_mul_host, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                    context(ℛ), context(params.ℛbig), UInt64(plain_modulus(params)), a, b, out, 1, C_NULL))
    else
        # default hooks (rlwe_she.jl:39-40): the component tensor over the ciphertext ring itself
        check(ccall((:tfb_ct_tensor_host, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                    context(ℛ), a, b, out, 1, C_NULL))
    end
    ntuple(k -> element(RingElement{ℛ}, proto, view(out, :, :, k)), 3)
end

# ---- device buffers: grown on demand, one per context and role (no malloc / free per call) ----
const DEVBUF = Dict{Tuple{Ptr{Cvoid},Symbol},Tuple{Ptr{Cvoid},Int}}()
function devbuf(ctx::Ptr{Cvoid}, tag::Symbol, bytes::Int)
    cur = get(DEVBUF, (ctx, tag), (C_NULL, 0))
    cur[2] >= bytes && return cur[1]
    cur[1] != C_NULL && check(ccall((:tfb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, cur[1]))
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tfb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx, bytes, p))
    DEVBUF[(ctx, tag)] = (p[], bytes)
    p[]
end
h2d(ctx, dst, src::Array{UInt64}) = check(ccall((:tfb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Csize_t, Ptr{Cvoid}), ctx, dst, src, sizeof(src), C_NULL))
d2h(ctx, dst::Array{UInt64}, src) = check(ccall((:tfb_memcpy_d2h, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), ctx, dst, src, sizeof(dst), C_NULL))
sync(ctx) = check(ccall((:tfb_sync, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, C_NULL))

# ---- keyswitch: rlwe_she.jl:315-349 ----
# The evaluation key is uploaded ONCE per (key, ciphertext level) in the NTT domain, [D][2][L'][N] with component 1 = mask,
# 2 = masked -- what the reference caches in key.mask.dual / key.masked.dual after first use (pow2_cyc_rings.jl:132-138) --
# restricted to the residues downswitch_keyelement selects for that level (modulusraising.jl:43-49: [1:l; special];
# crt.jl:238-244: 1:l).  `raised` follows the PARAMETER TYPE, not a comparison of rings.
struct DeviceKey
    ctx::Ptr{Cvoid}          # ring of the ciphertext
    ext::Ptr{Cvoid}          # ciphertext primes + special prime (ModulusRaised) or C_NULL
    ptr::Ptr{Cvoid}          # device buffer [D][2][L'][N]
    D::Int
    w::Int                   # relin_window (0 = CRT digits)
end
const DEVICE_KEYS = IdDict{Any,Dict{Int,DeviceKey}}()

function DeviceKey(ek::KeySwitchKey, ℛ::NegacyclicRing)
    raised = ek.params isa ModulusRaised
    ℛkey = ToyFHE.NTT.ring(ek.key[1].mask)
    l, lkey = nprimes(ℛ), nprimes(ℛkey)
    which = raised ? [1:l; lkey] : collect(1:l)
    ℛsel = ToyFHE.crtselect(ℛkey, which)
    N, Lsel, D = degree(ℛ), length(which), length(ek.key)
    host = Array{UInt64,4}(undef, N, Lsel, 2, D)
    tmp = Array{UInt64,3}(undef, N, Lsel, 1)
    for (k, kc) in enumerate(ek.key)
        host[:, :, 1, k] = pack!(tmp, 1, coeffs_dual(ToyFHE.crtselect(kc.mask, which)).parent)[:, :, 1]
        host[:, :, 2, k] = pack!(tmp, 1, coeffs_dual(ToyFHE.crtselect(kc.masked, which)).parent)[:, :, 1]
    end
    kctx = context(ℛsel)
    dev = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tfb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), kctx, sizeof(host), dev))
    check(ccall((:tfb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Csize_t, Ptr{Cvoid}), kctx, dev[], host, sizeof(host), C_NULL))
    sync(kctx)
    DeviceKey(context(ℛ), raised ? kctx : C_NULL, dev[], D, ToyFHE.relin_window(ek.params))
end
device_key(ek::KeySwitchKey, ℛ) = get!(() -> DeviceKey(ek, ℛ), get!(() -> Dict{Int,DeviceKey}(), DEVICE_KEYS, ek), nprimes(ℛ))

function ToyFHE.keyswitch(ek::KeySwitchKey, c::CipherText{Enc,P,T,NC}) where {Enc,P,ℛ,T<:RingElement{ℛ,<:CRTEncoded},NC}
    @assert NC in (2, 3)
    dk = device_key(ek, ℛ)
    proto = coeffs_primal(c[1])
    N, L = length(proto), nprimes(ℛ)
    host = staging(N, L, NC, :ks_in)
    for j in 1:NC
        pack!(host, j, coeffs_primal(c[j]).parent)
    end
    out = staging(N, L, 2, :ks_out)
    din, dout = devbuf(dk.ctx, :ks_in, sizeof(host)), devbuf(dk.ctx, :ks_out, sizeof(out))
    h2d(dk.ctx, din, host)
    check(ccall((:tfb_keyswitch, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt64, Ptr{Cvoid}),
                dk.ctx, dk.ext, dk.w, dk.ptr, dk.D, din, NC, dout, 1, C_NULL))
    d2h(dk.ctx, out, dout)
    sync(dk.ctx)
    CipherText{Enc}(c.params, ntuple(k -> element(RingElement{ℛ}, proto, view(out, :, :, k)), 2))
end

# ---- Galois automorphism: pow2_cyc_rings.jl:321-329 (rotate = keyswitch(gk, apply_galois_element(c, g)), rlwe_she.jl:355-359) ----
function NTT.apply_galois_element(re::RingElement{ℛ,Field}, galois_element::Integer) where {ℛ,Field<:CRTEncoded}
    ctx = context(ℛ)
    proto = coeffs_primal(re)
    N, L = length(proto), nprimes(ℛ)
    host = pack!(staging(N, L, 1, :gal), 1, proto.parent)
    din, dout = devbuf(ctx, :gal_in, sizeof(host)), devbuf(ctx, :gal_out, sizeof(host))
    h2d(ctx, din, host)
    check(ccall((:tfb_galois, LIB), Cint, (Ptr{Cvoid}, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{Cvoid}),
                ctx, UInt64(galois_element), din, dout, L, C_NULL))
    d2h(ctx, host, dout)
    sync(ctx)
    element(RingElement{ℛ}, proto, view(host, :, :, 1))
end

# ---- BFV plaintext maps (bfv.jl:21-29): Delta * m and mod(divround(SignedMod(x), Delta), t), exact on the device ----
limbs(x::Integer) = (n = cld(max(ndigits(x, base=2), 1), 64); UInt64[UInt64((x >> (64 * (i - 1))) & typemax(UInt64)) for i in 1:n])

function bfv_encode(ℛ, t::Integer, Δ::Integer, m::Vector{UInt64})
    d = limbs(Δ)
    out = Matrix{UInt64}(undef, length(m), nprimes(ℛ))
    check(ccall((:tfb_bfv_encode_host, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ptr{UInt64}, UInt32, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), UInt64(t), d, length(d), m, out, 1, C_NULL))
    out
end

function bfv_decode(ℛ, t::Integer, Δ::Integer, b::RingElement)
    d = limbs(Δ)
    p = coeffs_primal(b)
    buf = pack!(staging(length(p), nprimes(ℛ), 1, :dec), 1, p.parent)
    out = Vector{UInt64}(undef, length(p))
    check(ccall((:tfb_bfv_decode_host, LIB), Cint,
                (Ptr{Cvoid}, UInt64, Ptr{UInt64}, UInt32, Ptr{UInt64}, Ptr{UInt64}, UInt64, Ptr{Cvoid}),
                context(ℛ), UInt64(t), d, length(d), buf, out, 1, C_NULL))
    out
end

end # module
