"""sdr_b200 -- B200-native (sm_100a) streaming-FIR hot path of adamwalker/sdr behind the reference's own plugin API.

Python here is only the host-side mirror of the reference's Haskell interface (records, constructors, Pipes) over
the C ABI in ``include/sdr_b200.h``; all compute is in ``sdr_b200/lib/libsdr_b200.so`` (hand-written CUDA).
"""
from ._lib import (SDR_ARITH_EXACT, SDR_ARITH_FAST, SDR_DEVICE, SDR_DEVICE_HELD, SDR_HOST, SDR_HOST_PINNED, SdrError, LIB_PATH)  # noqa: F401
from .device import Context, DeviceBuffer, Event, PinnedArray, device_count, featureSelect, hasCUDA, has_cuda  # noqa: F401
from .filter import (Decimator, Filter, NativePipe, Resampler, cudaDecimatorC, cudaDecimatorR, cudaDecimatorSymR,  # noqa: F401
                     cudaFilterC, cudaFilterR, cudaFilterSymR, cudaResamplerC, cudaResamplerR, default_context,
                     firDecimator, firFilter, firResampler, pipeFirDecimator, pipeFirFilter, pipeFirResampler)
from .util import (complexFloatToInterleavedIQSigned2048, dcBlocker, dcBlockingFilter, fmDemod, pipeDcBlocker, fmDemodVec,  # noqa: F401
                   interleavedIQSigned2048ToFloat, interleavedIQUnsignedByteToFloat, pipeConvertU8, pipeFmDemod, pipeFmFrontEnd, pipeFmLowRate, pipeU8Decimator,
                   pipeScale, scaleFast)
from .filterdesign import windowed_sinc_taps  # noqa: F401
from . import multigpu, serialize  # noqa: F401
