"""TEST INFRASTRUCTURE ONLY -- Python restatement of the reference's Haskell host side for the FIR hot path.

Restates, line by line, the parts of ``hs_sources/SDR/Filter.hs`` and ``hs_sources/SDR/FilterInternal.hs`` that
sit between a Pipes stream of buffers and the C kernels:

* the plugin records ``Filter`` / ``Decimator`` / ``Resampler``            (Filter.hs:116-144)
* their constructors incl. coefficient padding / duplication / symmetry    (Filter.hs:163-502)
* ``prepareCoeffs`` polyphase table builder                                (FilterInternal.hs:277-319)
* the three streaming state machines ``firFilter`` / ``firDecimator`` / ``firResampler`` with their
  ``simple`` / ``crossover`` states and the output re-blocking ``advanceOutBuf``   (Filter.hs:504-727)
* ``fmDemod``                                                              (Demod.hs:38-46)

Kernel work is delegated to ``oracle.port()`` (our C restatement) or ``oracle.ref()`` (the compiled reference C);
cross-buffer kernels always use the port (they are Haskell in the reference, FilterInternal.hs:398-423).

Parity unpinned for the state machines themselves: the reference has no test or fixture for them
(tests/TestSuite.hs never runs a Pipe) and no GHC exists in this image; they are pinned only indirectly, by
``tests/test_oracle_pipes.py`` checking them against the closed-form flat-stream model.
"""
from dataclasses import dataclass
from typing import Any, Callable, Iterable, Iterator, List, Tuple

import numpy as np

import oracle
from oracle import V_AVX, V_AVX2, V_AVXSYM, V_SCALAR, V_SSE, V_SSE2, V_SSESYM


def round_up(num: int, div: int) -> int:
    """roundUp FilterInternal.hs:287-288"""
    return ((num + div - 1) // div) * div


def quot_up(q: int, d: int) -> int:
    """quotUp Filter.hs:675"""
    return (q + (d - 1)) // d


def duplicate(xs):
    """duplicate Filter.hs:146-148"""
    return np.repeat(np.asarray(xs, np.float32), 2)


# ---------------------------------------------------------------------------------------------------------------
# kernel back ends
# ---------------------------------------------------------------------------------------------------------------

_REF_NAME = {
    ("filter", False): {V_SCALAR: "filterRR", V_SSE: "filterSSERR", V_AVX: "filterAVXRR",
                        V_SSESYM: "filterSSESymmetricRR", V_AVXSYM: "filterAVXSymmetricRR"},
    ("filter", True): {V_SCALAR: "filterRC", V_SSE: "filterSSERC", V_AVX: "filterAVXRC", V_SSE2: "filterSSERC2",
                       V_AVX2: "filterAVXRC2", V_SSESYM: "filterSSESymmetricRC", V_AVXSYM: "filterAVXSymmetricRC"},
    ("decimate", False): {V_SCALAR: "decimateRR", V_SSE: "decimateSSERR", V_AVX: "decimateAVXRR",
                          V_SSESYM: "decimateSSESymmetricRR", V_AVXSYM: "decimateAVXSymmetricRR"},
    ("decimate", True): {V_SCALAR: "decimateRC", V_SSE: "decimateSSERC", V_AVX: "decimateAVXRC",
                         V_SSE2: "decimateSSERC2", V_AVX2: "decimateAVXRC2", V_SSESYM: "decimateSSESymmetricRC",
                         V_AVXSYM: "decimateAVXSymmetricRC"},
    ("resample", False): {V_SCALAR: "resample2RR", V_SSE: "resampleSSERR", V_AVX: "resampleAVXRR"},
    ("resample", True): {V_SCALAR: "resample2RC", V_SSE2: "resampleSSERC", V_AVX2: "resampleAVXRC"},
}


class Kernels:
    """Uniform call surface over the port (backend='port') or the compiled reference (backend='ref')."""

    def __init__(self, backend="port"):
        self.backend = backend
        self.port = oracle.port()
        self.ref = oracle.ref() if backend == "ref" else None
        if backend == "ref" and self.ref is None:
            raise RuntimeError("oracle/_ref/libsdrref.so not built")

    def decimate(self, variant, num, factor, coeffs, x, cplx):
        if self.ref is not None:
            return self.ref.decimate(_REF_NAME[("decimate", cplx)][variant], num, factor, coeffs, x)
        return self.port.decimate(variant, num, factor, coeffs, x, cplx)

    def filter(self, variant, num, coeffs, x, cplx):
        if self.ref is not None:
            return self.ref.filter(_REF_NAME[("filter", cplx)][variant], num, coeffs, x)
        return self.port.filter(variant, num, coeffs, x, cplx)

    def resample(self, variant, num, num_coeffs, group, increments, groups, x, cplx):
        if self.ref is not None:
            return self.ref.resample(_REF_NAME[("resample", cplx)][variant], num, num_coeffs, group, increments,
                                     groups, x)
        return self.port.resample_n(variant, num, num_coeffs, group, increments, groups, x, cplx)


# ---------------------------------------------------------------------------------------------------------------
# plugin records (Filter.hs:116-144)
# ---------------------------------------------------------------------------------------------------------------

@dataclass
class Filter:
    numCoeffsF: int
    filterOne: Callable      # count -> bufIn -> out[count]
    filterCross: Callable    # count -> bufLast -> bufNext -> out[count]
    cplx: bool = False


@dataclass
class Decimator:
    numCoeffsD: int
    decimationD: int
    decimateOne: Callable
    decimateCross: Callable
    cplx: bool = False


@dataclass
class Resampler:
    numCoeffsR: int
    decimationR: int
    interpolationR: int
    startDat: Any
    resampleOne: Callable    # dat -> count -> bufIn -> (out[count], (dat', endOffset))
    resampleCross: Callable  # dat -> count -> bufLast -> bufNext -> (out[count], (dat', endOffset))
    cplx: bool = False


_SIZE_MULT_R = {V_SCALAR: 1, V_SSE: 4, V_AVX: 8}   # Filter.hs:180,185,190 / :296,302,308
_SIZE_MULT_C = {V_SCALAR: 1, V_SSE: 2, V_AVX: 4}   # Filter.hs:216,221,226 / :337,343,349


def mk_filter(variant, coeffs, k: Kernels = None) -> Filter:
    """mkFilter Filter.hs:163-175 (fastFilterCR / SSER / AVXR)"""
    k = k or Kernels()
    l = len(coeffs)
    n = round_up(l, _SIZE_MULT_R[variant])
    v = np.concatenate([np.asarray(coeffs, np.float32), np.zeros(n - l, np.float32)])
    return Filter(n,
                  lambda count, buf: k.filter(variant, count, v, buf, False),
                  lambda count, last, nxt: k.port.decimate_cross(1, v, count, last, nxt, False))


def mk_filter_c(variant, coeffs, k: Kernels = None) -> Filter:
    """mkFilterC Filter.hs:198-211.  NOTE the reference calls `roundUp sizeMultiple l` with the arguments swapped
    (:204), which for l >= sizeMultiple returns l: complex filters are never padded.  Restated as is."""
    k = k or Kernels()
    l = len(coeffs)
    n = round_up(_SIZE_MULT_C[variant], l)
    padded = np.concatenate([np.asarray(coeffs, np.float32), np.zeros(n - l, np.float32)])
    # Reference quirk restated as is: the duplicated vector is handed to EVERY variant, including the scalar
    # filterCRC of fastFilterCC (:206,209,216), whose dotprod_C expects plain taps -- so that (never selected on an
    # AVX/SSE4.2 host, CPUID.hs:100-104) path computes a different, 2T-long filter.  Not replicated by the product.
    v = duplicate(padded)
    return Filter(n,
                  lambda count, buf: k.filter(variant, count, v, buf, True),
                  lambda count, last, nxt: k.port.decimate_cross(1, padded, count, last, nxt, True),
                  cplx=True)


def mk_filter_sym_r(variant, half, k: Kernels = None) -> Filter:
    """mkFilterSymR Filter.hs:234-245; variant in (V_SSESYM, V_AVXSYM); `half` = first half of the taps."""
    k = k or Kernels()
    v = np.asarray(half, np.float32)
    v2 = np.concatenate([v, v[::-1]])
    return Filter(2 * len(v),
                  lambda count, buf: k.filter(variant, count, v, buf, False),
                  lambda count, last, nxt: k.port.decimate_cross(1, v2, count, last, nxt, False))


def mk_decimator(variant, factor, coeffs, k: Kernels = None) -> Decimator:
    """mkDecimator Filter.hs:277-290"""
    k = k or Kernels()
    l = len(coeffs)
    n = round_up(l, _SIZE_MULT_R[variant])
    v = np.concatenate([np.asarray(coeffs, np.float32), np.zeros(n - l, np.float32)])
    return Decimator(n, factor,
                     lambda count, buf: k.decimate(variant, count, factor, v, buf, False),
                     lambda count, last, nxt: k.port.decimate_cross(factor, v, count, last, nxt, False))


def mk_decimator_c(variant, factor, coeffs, k: Kernels = None) -> Decimator:
    """mkDecimatorC Filter.hs:317-331 (fastDecimatorCC / SSEC / AVXC); pads to x1 / x2 / x4 and, for the SIMD
    variants, duplicates each tap (:326)."""
    k = k or Kernels()
    l = len(coeffs)
    n = round_up(l, _SIZE_MULT_C[variant])
    padded = np.concatenate([np.asarray(coeffs, np.float32), np.zeros(n - l, np.float32)])
    v = duplicate(padded)   # for every variant, scalar decimateCRC included -- same quirk as mk_filter_c
    return Decimator(n, factor,
                     lambda count, buf: k.decimate(variant, count, factor, v, buf, True),
                     lambda count, last, nxt: k.port.decimate_cross(factor, padded, count, last, nxt, True),
                     cplx=True)


def mk_decimator_sym_r(variant, factor, half, k: Kernels = None) -> Decimator:
    """mkDecimatorSymR Filter.hs:358-371"""
    k = k or Kernels()
    v = np.asarray(half, np.float32)
    v2 = np.concatenate([v, v[::-1]])
    return Decimator(2 * len(v), factor,
                     lambda count, buf: k.decimate(variant, count, factor, v, buf, False),
                     lambda count, last, nxt: k.port.decimate_cross(factor, v2, count, last, nxt, False))


def stride_list(s, xs):
    """strideList FilterInternal.hs:280-285"""
    return list(xs[::s])


def prepare_coeffs(n, interpolation, decimation, coeffs):
    """prepareCoeffs FilterInternal.hs:297-319 -> (numCoeffs, increments, groups[numGroups, roundUp numCoeffs n])."""
    coeffs = list(np.asarray(coeffs, np.float32))
    dats = []
    offset = 0
    while True:
        q, r = divmod(decimation - offset - 1, interpolation)
        dats.append((q + 1, stride_list(interpolation, coeffs[offset:])))
        offset = interpolation - 1 - r
        if offset == 0:        # func' 0 = []
            break
    num_coeffs = max(len(g) for _, g in dats)
    width = round_up(num_coeffs, n)
    groups = np.zeros((len(dats), width), np.float32)
    for i, (_, g) in enumerate(dats):
        groups[i, :len(g)] = g
    increments = [inc for inc, _ in dats]
    return num_coeffs, increments, groups


_SIZE_MULT_RS = {V_SCALAR: 1, V_SSE: 4, V_AVX: 8, V_SSE2: 4, V_AVX2: 8}   # Filter.hs:451,458,465,480,487,494


def mk_resampler(variant, interpolation, decimation, coeffs, cplx=False, k: Kernels = None) -> Resampler:
    """Filter.mkResampler / mkResamplerC Filter.hs:408-444 over FilterInternal.mkResampler(C) :335-373.
    dat = (group, offset); resampleOne ignores the incoming offset and recomputes it from the group C returns
    (func1, Filter.hs:423)."""
    k = k or Kernels()
    sm = _SIZE_MULT_RS[variant]
    v = np.asarray(coeffs, np.float32)
    num_coeffs, increments, groups = prepare_coeffs(sm, interpolation, decimation, v)

    def func1(group):
        offset = interpolation - 1 - ((interpolation + group * decimation - 1) % interpolation)
        return (group, offset), offset

    def resample_one(dat, count, buf):
        out, group = k.resample(variant, count, num_coeffs, dat[0], increments, groups, buf, cplx)
        return out, func1(group)

    def resample_cross(dat, count, last, nxt):
        group, offset = dat
        out, offset2 = k.port.resample_cross(interpolation, decimation, v, offset, count, last, nxt, cplx)
        return out, (((group + count) % interpolation, offset2), offset2)

    return Resampler(round_up(len(v), interpolation * sm), decimation, interpolation, (0, 0), resample_one,
                     resample_cross, cplx=cplx)


# ---------------------------------------------------------------------------------------------------------------
# output re-blocking (Filter.hs:504-523)
# ---------------------------------------------------------------------------------------------------------------

class _OutBuf:
    def __init__(self, size, dtype):
        self.buf = np.zeros(size, dtype)   # VGM.new
        self.offset = 0

    def space(self):
        return len(self.buf) - self.offset


def _advance(block_size_out, ob: _OutBuf, count, dtype, emit: List):
    """advanceOutBuf Filter.hs:516-523"""
    if count == ob.space():
        emit.append(ob.buf)                 # unsafeFreeze + yield
        return _OutBuf(block_size_out, dtype)
    ob.offset += count
    return ob


def _assert(loc, cond):
    """assert Filter.hs:525-527"""
    if not cond:
        raise RuntimeError(loc)


# ---------------------------------------------------------------------------------------------------------------
# the three Pipes (generators: pull input vectors from `src`, yield output vectors of exactly blockSizeOut)
# ---------------------------------------------------------------------------------------------------------------

def fir_filter(f: Filter, block_size_out: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firFilter Filter.hs:532-569"""
    it = iter(src)
    dtype = np.complex64 if f.cplx else np.float32
    try:
        buf_in = next(it)
    except StopIteration:
        return
    ob = _OutBuf(block_size_out, dtype)
    state, buf_last = "simple", None
    while True:
        emit: List[np.ndarray] = []
        if state == "simple":
            _assert("filter 1", len(buf_in) >= f.numCoeffsF)
            count = min(len(buf_in) - f.numCoeffsF + 1, ob.space())
            ob.buf[ob.offset:ob.offset + count] = f.filterOne(count, buf_in)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            buf_in = buf_in[count:]
            yield from emit
            if len(buf_in) < f.numCoeffsF:
                try:
                    nxt = next(it)
                except StopIteration:
                    return
                buf_last, buf_in, state = buf_in, nxt, "cross"
        else:
            _assert("filter 2", len(buf_last) < f.numCoeffsF)
            _assert("filter 3", len(buf_last) > 0)
            count = min(len(buf_last), ob.space())
            ob.buf[ob.offset:ob.offset + count] = f.filterCross(count, buf_last, buf_in)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            yield from emit
            if len(buf_last) == count:
                state = "simple"
            else:
                buf_last = buf_last[count:]


def fir_decimator(d: Decimator, block_size_out: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firDecimator Filter.hs:574-611"""
    it = iter(src)
    dtype = np.complex64 if d.cplx else np.float32
    D = d.decimationD
    try:
        buf_in = next(it)
    except StopIteration:
        return
    ob = _OutBuf(block_size_out, dtype)
    state, buf_last = "simple", None
    while True:
        emit: List[np.ndarray] = []
        if state == "simple":
            _assert("decimate 1", len(buf_in) >= d.numCoeffsD)
            count = min((len(buf_in) - d.numCoeffsD) // D + 1, ob.space())
            ob.buf[ob.offset:ob.offset + count] = d.decimateOne(count, buf_in)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            buf_in = buf_in[count * D:]
            yield from emit
            if len(buf_in) < d.numCoeffsD:
                try:
                    nxt = next(it)
                except StopIteration:
                    return
                buf_last, buf_in, state = buf_in, nxt, "cross"
        else:
            _assert("decimate 2", len(buf_last) < d.numCoeffsD)
            _assert("decimate 3", len(buf_last) > 0)
            count = min(quot_up(len(buf_last), D), ob.space())
            ob.buf[ob.offset:ob.offset + count] = d.decimateCross(count, buf_last, buf_in)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            yield from emit
            if len(buf_last) <= count * D:
                buf_in = buf_in[count * D - len(buf_last):]
                state = "simple"
            else:
                buf_last = buf_last[count * D:]


def fir_resampler(r: Resampler, block_size_out: int, src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """firResampler Filter.hs:679-727"""
    it = iter(src)
    dtype = np.complex64 if r.cplx else np.float32
    L, M, T = r.interpolationR, r.decimationR, r.numCoeffsR
    try:
        buf_in = next(it)
    except StopIteration:
        return
    ob = _OutBuf(block_size_out, dtype)
    state, buf_last = "simple", None
    dat, filter_offset = r.startDat, 0
    while True:
        emit: List[np.ndarray] = []
        if state == "simple":
            _assert("resample 1", len(buf_in) * L >= T - filter_offset)
            count = min((len(buf_in) * L - T + filter_offset) // M + 1, ob.space())
            out, (dat, end_offset) = r.resampleOne(dat, count, buf_in)
            ob.buf[ob.offset:ob.offset + count] = out
            _assert("resample 2", (count * M + end_offset - filter_offset) % L == 0)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            used = quot_up(count * M - filter_offset, L)
            buf_in = buf_in[used:]
            filter_offset = end_offset
            yield from emit
            if len(buf_in) * L < T - end_offset:
                try:
                    nxt = next(it)
                except StopIteration:
                    return
                if len(buf_in) == 0:          # Filter.hs:708-710
                    buf_in = nxt
                else:
                    buf_last, buf_in, state = buf_in, nxt, "cross"
        else:
            _assert("resample 3", len(buf_last) * L < T - filter_offset)
            computable = quot_up(len(buf_last) * L + filter_offset, M)
            count = min(computable, ob.space())
            _assert("resample 4", count != 0)
            out, (dat, end_offset) = r.resampleCross(dat, count, buf_last, buf_in)
            ob.buf[ob.offset:ob.offset + count] = out
            _assert("resample 5", (count * M + end_offset - filter_offset) % L == 0)
            ob = _advance(block_size_out, ob, count, dtype, emit)
            used = quot_up(count * M - filter_offset, L)
            filter_offset = end_offset
            yield from emit
            if used >= len(buf_last):
                buf_in = buf_in[used - len(buf_last):]
                state = "simple"
            else:
                buf_last = buf_last[used:]


def fm_demod(src: Iterable[np.ndarray]) -> Iterator[np.ndarray]:
    """fmDemod Demod.hs:38-46"""
    last = 0j
    for dat in src:
        yield oracle.port().fm_demod(dat, last)
        last = complex(dat[-1])


# ---------------------------------------------------------------------------------------------------------------
# closed-form flat-stream model in float64 -- the size-independent property the state
# machines and every kernel are checked against.
# ---------------------------------------------------------------------------------------------------------------

def flat_decimate(x, taps, factor=1, num=None):
    """y[m] = sum_k c[k] x[m*D + k], valid mode, float64 (complex128 for complex input)."""
    x = np.asarray(x)
    c = np.asarray(taps, np.float64)
    T = len(c)
    n_out = (len(x) - T) // factor + 1 if len(x) >= T else 0
    if num is not None:
        n_out = min(n_out, num)
    acc = np.zeros(n_out, np.complex128 if np.iscomplexobj(x) else np.float64)
    xs = x.astype(acc.dtype)
    span = (n_out - 1) * factor + 1 if n_out > 0 else 0
    for k in range(T):
        acc += c[k] * xs[k:k + span:factor]
    return acc


def flat_resample(x, taps, interpolation, decimation, num=None, padded_taps=None):
    """y[k] = sum_l c[f_k + l L] x[i_k + l], f_k = (-k M) mod L, i_k = ceil(k M / L); float64.
    Output count for a finite stream uses the padded tap count T' (SURVEY.md 8a): (N*L - T') div M + 1."""
    x = np.asarray(x)
    c = np.asarray(taps, np.float64)
    L, M, T = interpolation, decimation, len(c)
    Tp = padded_taps or T
    n_out = (len(x) * L - Tp) // M + 1 if len(x) * L >= Tp else 0
    if num is not None:
        n_out = min(n_out, num)
    dtype = np.complex128 if np.iscomplexobj(x) else np.float64
    xs = np.concatenate([x.astype(dtype), np.zeros(T, dtype)])
    out = np.zeros(n_out, dtype)
    ks = np.arange(n_out)
    for ph in range(L):
        sel = ks[ks % L == ph]
        if len(sel) == 0:
            continue
        f = (-sel[0] * M) % L
        i0 = -((-sel * M) // L)          # ceil(k M / L)
        ph_taps = c[f::L]
        acc = np.zeros(len(sel), dtype)
        for l, t in enumerate(ph_taps):
            acc += t * xs[i0 + l]
        out[sel] = acc
    return out
