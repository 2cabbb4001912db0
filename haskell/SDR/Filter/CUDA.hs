{-# LANGUAGE ForeignFunctionInterface #-}
{-| B200 (CUDA) members of the @fast*@ constructor families of "SDR.Filter", over libsdr_b200's C ABI
    (include/sdr_b200.h).  They return the library's own plugin records, so 'SDR.Filter.firFilter',
    'SDR.Filter.firDecimator' and 'SDR.Filter.firResampler' and all Pipes glue are used unchanged.

    This module cannot be compiled in the build image (no GHC); it is the binding a maintainer would add, kept
    deliberately mechanical: one @foreign import@ per C entry point, one constructor per reference constructor.
    Compare hs_sources/SDR/FilterInternal.hs:80-249 (the imports it replaces) and hs_sources/SDR/Filter.hs:277-389.
-}
module SDR.Filter.CUDA (
    hasCUDA,
    cudaFilterR, cudaFilterC, cudaFilterSymR,
    cudaDecimatorR, cudaDecimatorC, cudaDecimatorSymR,
    cudaResamplerR, cudaResamplerC,
    cudaDecimatorOrFast,
    -- * Record handles for the device-resident stages of "SDR.Pipes.CUDA"
    FilterH, DecimatorH, ResamplerH, Ctx, defaultCtx,
    filterHandle, decimatorHandle, resamplerHandle
    ) where

import           Control.Monad                (unless, when)
import           Data.Complex
import qualified Data.Vector.Storable         as VS
import qualified Data.Vector.Storable.Mutable as VSM
import           Foreign.C.String
import           Foreign.C.Types
import           Foreign.ForeignPtr
import           Foreign.Marshal.Alloc
import           Foreign.Marshal.Utils        (with)
import           Foreign.Ptr
import           Foreign.Storable

import           SDR.CPUID                    (CPUInfo)
import           SDR.Filter                   (Decimator (..), Filter (..), Resampler (..), fastDecimatorC)

-- opaque handles ------------------------------------------------------------------------------------------------
data Ctx
data FilterH
data DecimatorH
data ResamplerH

-- sdr_resampler_dat_t { int group; int offset; }
data Dat = Dat !CInt !CInt
instance Storable Dat where
    sizeOf _    = 8
    alignment _ = 4
    peek p      = Dat <$> peekByteOff p 0 <*> peekByteOff p 4
    poke p (Dat g o) = pokeByteOff p 0 g >> pokeByteOff p 4 o

foreign import ccall unsafe "sdr_has_cuda"            c_hasCuda      :: IO CInt
foreign import ccall unsafe "sdr_last_error"          c_lastError    :: IO CString
foreign import ccall safe   "sdr_ctx_create"          c_ctxCreate    :: CInt -> Ptr (Ptr Ctx) -> IO CInt

foreign import ccall safe   "sdr_filter_create"       c_filterCreate    :: Ptr Ctx -> CInt -> Ptr CFloat -> CInt -> CInt -> Ptr (Ptr FilterH) -> IO CInt
foreign import ccall safe   "sdr_filter_create_sym"   c_filterCreateSym :: Ptr Ctx -> CInt -> Ptr CFloat -> CInt -> Ptr (Ptr FilterH) -> IO CInt
foreign import ccall unsafe "sdr_filter_num_coeffs"   c_filterNumCoeffs :: Ptr FilterH -> IO CInt
foreign import ccall safe   "sdr_filter_one"          c_filterOne       :: Ptr FilterH -> CInt -> Ptr a -> Ptr a -> CInt -> IO CInt
foreign import ccall safe   "sdr_filter_cross"        c_filterCross     :: Ptr FilterH -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> IO CInt
foreign import ccall unsafe "&sdr_filter_destroy"     p_filterDestroy   :: FunPtr (Ptr FilterH -> IO ())

foreign import ccall safe   "sdr_decimator_create"     c_decimatorCreate    :: Ptr Ctx -> CInt -> CInt -> Ptr CFloat -> CInt -> CInt -> Ptr (Ptr DecimatorH) -> IO CInt
foreign import ccall safe   "sdr_decimator_create_sym" c_decimatorCreateSym :: Ptr Ctx -> CInt -> CInt -> Ptr CFloat -> CInt -> Ptr (Ptr DecimatorH) -> IO CInt
foreign import ccall unsafe "sdr_decimator_num_coeffs" c_decimatorNumCoeffs :: Ptr DecimatorH -> IO CInt
foreign import ccall safe   "sdr_decimate_one"         c_decimateOne        :: Ptr DecimatorH -> CInt -> Ptr a -> Ptr a -> CInt -> IO CInt
foreign import ccall safe   "sdr_decimate_cross"       c_decimateCross      :: Ptr DecimatorH -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> IO CInt
foreign import ccall unsafe "&sdr_decimator_destroy"   p_decimatorDestroy   :: FunPtr (Ptr DecimatorH -> IO ())

foreign import ccall safe   "sdr_resampler_create"     c_resamplerCreate    :: Ptr Ctx -> CInt -> CInt -> CInt -> Ptr CFloat -> CInt -> CInt -> Ptr (Ptr ResamplerH) -> IO CInt
foreign import ccall unsafe "sdr_resampler_num_coeffs" c_resamplerNumCoeffs :: Ptr ResamplerH -> IO CInt
foreign import ccall safe   "sdr_resample_one"         c_resampleOne        :: Ptr ResamplerH -> Ptr Dat -> CInt -> Ptr a -> Ptr a -> CInt -> Ptr CInt -> IO CInt
foreign import ccall safe   "sdr_resample_cross"       c_resampleCross      :: Ptr ResamplerH -> Ptr Dat -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> Ptr CInt -> IO CInt
foreign import ccall unsafe "&sdr_resampler_destroy"   p_resamplerDestroy   :: FunPtr (Ptr ResamplerH -> IO ())

sdrHost :: CInt
sdrHost = 0

-- | the predicate that slots into 'SDR.CPUID.featureSelect' (CPUID.hs:100-104)
hasCUDA :: IO Bool
hasCUDA = (/= 0) <$> c_hasCuda

-- | non-zero status -> the library's message as an 'error', like the reference's own asserts (Filter.hs:525-527)
check :: IO CInt -> IO ()
check act = do
    st <- act
    unless (st == 0) $ c_lastError >>= peekCString >>= error

defaultCtx :: IO (Ptr Ctx)
defaultCtx = alloca $ \pp -> check (c_ctxCreate 0 pp) >> peek pp

withCoeffs :: [Float] -> (Ptr CFloat -> CInt -> IO b) -> IO b
withCoeffs cs f = VS.unsafeWith (VS.fromList cs) $ \p -> f (castPtr p) (fromIntegral (length cs))

-- Filters --------------------------------------------------------------------------------------------------------
mkCudaFilter :: (Storable a) => Bool -> Bool -> [Float] -> IO (Filter IO VS.Vector VS.MVector a)
mkCudaFilter cplx sym coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n ->
            if sym then check (c_filterCreateSym ctx (fromBool' cplx) pc n pp)
                   else check (c_filterCreate ctx (fromBool' cplx) pc n 1 pp)
        peek pp
    fp  <- newForeignPtr p_filterDestroy h
    n   <- fromIntegral <$> c_filterNumCoeffs h
    let one num inBuf outBuf = withForeignPtr fp $ \hp ->
            VS.unsafeWith inBuf $ \iPtr -> VSM.unsafeWith outBuf $ \oPtr ->
                check $ c_filterOne hp (fromIntegral num) iPtr oPtr sdrHost
        cross num lastBuf nextBuf outBuf = withForeignPtr fp $ \hp ->
            VS.unsafeWith lastBuf $ \lPtr -> VS.unsafeWith nextBuf $ \nPtr -> VSM.unsafeWith outBuf $ \oPtr ->
                check $ c_filterCross hp (fromIntegral num) lPtr (fromIntegral (VS.length lastBuf)) nPtr
                                      (fromIntegral (VS.length nextBuf)) oPtr sdrHost
    return $ Filter n one cross
  where fromBool' b = if b then 1 else 0

-- | CUDA member of the 'SDR.Filter.fastFilterR' family (Filter.hs:193-196)
cudaFilterR :: [Float] -> IO (Filter IO VS.Vector VS.MVector Float)
cudaFilterR = mkCudaFilter False False

-- | 'SDR.Filter.fastFilterC' (Filter.hs:229-232)
cudaFilterC :: [Float] -> IO (Filter IO VS.Vector VS.MVector (Complex Float))
cudaFilterC = mkCudaFilter True False

-- | 'SDR.Filter.fastFilterSymR' (Filter.hs:258-261): pass the first half of the taps
cudaFilterSymR :: [Float] -> IO (Filter IO VS.Vector VS.MVector Float)
cudaFilterSymR = mkCudaFilter False True

-- Decimators -----------------------------------------------------------------------------------------------------
mkCudaDecimator :: (Storable a) => Bool -> Bool -> Int -> [Float] -> IO (Decimator IO VS.Vector VS.MVector a)
mkCudaDecimator cplx sym factor coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n ->
            if sym then check (c_decimatorCreateSym ctx (b cplx) (fromIntegral factor) pc n pp)
                   else check (c_decimatorCreate ctx (b cplx) (fromIntegral factor) pc n 1 pp)
        peek pp
    fp  <- newForeignPtr p_decimatorDestroy h
    n   <- fromIntegral <$> c_decimatorNumCoeffs h
    let one num inBuf outBuf = withForeignPtr fp $ \hp ->
            VS.unsafeWith inBuf $ \iPtr -> VSM.unsafeWith outBuf $ \oPtr ->
                check $ c_decimateOne hp (fromIntegral num) iPtr oPtr sdrHost
        cross num lastBuf nextBuf outBuf = withForeignPtr fp $ \hp ->
            VS.unsafeWith lastBuf $ \lPtr -> VS.unsafeWith nextBuf $ \nPtr -> VSM.unsafeWith outBuf $ \oPtr ->
                check $ c_decimateCross hp (fromIntegral num) lPtr (fromIntegral (VS.length lastBuf)) nPtr
                                        (fromIntegral (VS.length nextBuf)) oPtr sdrHost
    return $ Decimator n factor one cross
  where b x = if x then 1 else 0

-- | 'SDR.Filter.fastDecimatorR' (Filter.hs:311-315)
cudaDecimatorR :: Int -> [Float] -> IO (Decimator IO VS.Vector VS.MVector Float)
cudaDecimatorR = mkCudaDecimator False False

-- | 'SDR.Filter.fastDecimatorC' (Filter.hs:352-356) -- the headline path
cudaDecimatorC :: Int -> [Float] -> IO (Decimator IO VS.Vector VS.MVector (Complex Float))
cudaDecimatorC = mkCudaDecimator True False

-- | 'SDR.Filter.fastDecimatorSymR' (Filter.hs:385-389)
cudaDecimatorSymR :: Int -> [Float] -> IO (Decimator IO VS.Vector VS.MVector Float)
cudaDecimatorSymR = mkCudaDecimator False True

-- | featureSelect with CUDA in front: what examples/fm/fm.hs:30 would call instead of fastDecimatorC
cudaDecimatorOrFast :: CPUInfo -> Int -> [Float] -> IO (Decimator IO VS.Vector VS.MVector (Complex Float))
cudaDecimatorOrFast info factor coeffs = do
    cuda <- hasCUDA
    if cuda then cudaDecimatorC factor coeffs else fastDecimatorC info factor coeffs

-- Resamplers -----------------------------------------------------------------------------------------------------
mkCudaResampler :: (Storable a) => Bool -> Int -> Int -> [Float] -> IO (Resampler IO VS.Vector VS.MVector a)
mkCudaResampler cplx interpolation decimation coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n ->
            check (c_resamplerCreate ctx (if cplx then 1 else 0) (fromIntegral interpolation) (fromIntegral decimation) pc n 1 pp)
        peek pp
    fp  <- newForeignPtr p_resamplerDestroy h
    n   <- fromIntegral <$> c_resamplerNumCoeffs h
    -- the record's existential state is (group, offset), exactly as in Filter.hs:424
    let one (g, o) num inBuf outBuf = withForeignPtr fp $ \hp ->
            with (Dat (fromIntegral g) (fromIntegral o)) $ \pd -> alloca $ \pe ->
            VS.unsafeWith inBuf $ \iPtr -> VSM.unsafeWith outBuf $ \oPtr -> do
                check $ c_resampleOne hp pd (fromIntegral num) iPtr oPtr sdrHost pe
                Dat g' o' <- peek pd
                e         <- peek pe
                return ((fromIntegral g', fromIntegral o'), fromIntegral e)
        cross (g, o) num lastBuf nextBuf outBuf = withForeignPtr fp $ \hp ->
            with (Dat (fromIntegral g) (fromIntegral o)) $ \pd -> alloca $ \pe ->
            VS.unsafeWith lastBuf $ \lPtr -> VS.unsafeWith nextBuf $ \nPtr -> VSM.unsafeWith outBuf $ \oPtr -> do
                check $ c_resampleCross hp pd (fromIntegral num) lPtr (fromIntegral (VS.length lastBuf)) nPtr
                                        (fromIntegral (VS.length nextBuf)) oPtr sdrHost pe
                Dat g' o' <- peek pd
                e         <- peek pe
                return ((fromIntegral g', fromIntegral o'), fromIntegral e)
    return $ Resampler n decimation interpolation (0 :: Int, 0 :: Int) one cross

-- | 'SDR.Filter.fastResamplerR' (Filter.hs:468-473)
cudaResamplerR :: Int -> Int -> [Float] -> IO (Resampler IO VS.Vector VS.MVector Float)
cudaResamplerR = mkCudaResampler False

-- | 'SDR.Filter.fastResamplerC' (Filter.hs:497-502)
cudaResamplerC :: Int -> Int -> [Float] -> IO (Resampler IO VS.Vector VS.MVector (Complex Float))
cudaResamplerC = mkCudaResampler True

-- Handles for "SDR.Pipes.CUDA" ---------------------------------------------------------------------------------------
-- The device-resident stages (sdr_pipe_fir_decimator ...) take the plugin record by handle and keep referring to it,
-- so they receive the ForeignPtr itself (the stage stores it: the record cannot be finalized before the stage).

-- | the handle of a 'cudaFilterR' / 'cudaFilterC' / 'cudaFilterSymR' style record: complex data?, symmetric half taps?
filterHandle :: Bool -> Bool -> [Float] -> IO (ForeignPtr FilterH)
filterHandle cplx sym coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n ->
            if sym then check (c_filterCreateSym ctx (if cplx then 1 else 0) pc n pp)
                   else check (c_filterCreate ctx (if cplx then 1 else 0) pc n 1 pp)
        peek pp
    newForeignPtr p_filterDestroy h

-- | the handle of a 'cudaDecimatorC' style record
decimatorHandle :: Bool -> Int -> [Float] -> IO (ForeignPtr DecimatorH)
decimatorHandle cplx factor coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n -> check (c_decimatorCreate ctx (if cplx then 1 else 0) (fromIntegral factor) pc n 1 pp)
        peek pp
    newForeignPtr p_decimatorDestroy h

-- | the handle of a 'cudaResamplerR' style record
resamplerHandle :: Bool -> Int -> Int -> [Float] -> IO (ForeignPtr ResamplerH)
resamplerHandle cplx interpolation decimation coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        withCoeffs coeffs $ \pc n ->
            check (c_resamplerCreate ctx (if cplx then 1 else 0) (fromIntegral interpolation) (fromIntegral decimation) pc n 1 pp)
        peek pp
    newForeignPtr p_resamplerDestroy h
