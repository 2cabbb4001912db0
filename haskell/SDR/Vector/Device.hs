{-# LANGUAGE ForeignFunctionInterface, MultiParamTypeClasses, TypeFamilies, ScopedTypeVariables, FlexibleInstances #-}
{-| A device-resident vector type with "Data.Vector.Generic" instances, so that the UNMODIFIED polymorphic pipes of
    "SDR.Filter" -- @firDecimator :: (PrimMonad m, Functor m, VG.Vector v a) => Decimator m v (VG.Mutable v) a -> Int ->
    Pipe (v a) (v a) m ()@ (hs_sources/SDR/Filter.hs:574-577) and its siblings -- carry buffers that live in HBM.

    The pipes touch their buffers only through @VG.length@, @VG.drop@, @VGM.unsafeDrop@, @VGM.new@ and
    @VG.unsafeFreeze@ (Filter.hs:513, 519, 546-550); all five are O(1) pointer arithmetic or one allocation here, exactly
    as for storable vectors, and the record closures built by 'deviceDecimatorC' call libsdr_b200's layer 2 with
    @SDR_DEVICE@ pointers (include/sdr_b200.h: sdr_decimate_one / sdr_decimate_cross).  Element access from Haskell
    ('basicUnsafeIndexM') copies one element over PCIe: fine for a debugger, never used by the pipes.

    Source only: GHC is not in the build image, so this module has never been compiled.  It is the last piece of the
    binding a maintainer would add (SURVEY.md section 7 step 8); everything it calls is exercised through the Python
    host mirror with device pointers (sdr_b200/filter.py, tests/test_gpu_parity.py).
-}
module SDR.Vector.Device (
    DVector, DMVector,
    fromStorable, toStorable,
    deviceDecimatorC
    ) where

import           Control.Monad                (unless)
import           Control.Monad.Primitive      (unsafePrimToPrim)
import           Data.Complex
import qualified Data.Vector.Generic          as VG
import qualified Data.Vector.Generic.Mutable  as VGM
import qualified Data.Vector.Storable         as VS
import qualified Data.Vector.Storable.Mutable as VSM
import           Foreign.C.String
import           Foreign.C.Types
import qualified Foreign.Concurrent           as FC
import           Foreign.ForeignPtr
import           Foreign.Marshal.Alloc
import           Foreign.Ptr
import           Foreign.Storable
import           System.IO.Unsafe             (unsafePerformIO)

import           SDR.Filter                   (Decimator (..))

data Ctx
data DecimatorH

foreign import ccall unsafe "sdr_last_error"           c_lastError  :: IO CString
foreign import ccall safe   "sdr_ctx_create"           c_ctxCreate  :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe   "sdr_ctx_sync"             c_ctxSync    :: Ptr Ctx -> IO CInt
foreign import ccall safe   "sdr_dev_alloc"            c_devAlloc   :: Ptr Ctx -> CSize -> Ptr (Ptr a) -> IO CInt
foreign import ccall safe   "sdr_dev_free"             c_devFree    :: Ptr Ctx -> Ptr a -> IO CInt
foreign import ccall safe   "sdr_memcpy_h2d"           c_h2d        :: Ptr Ctx -> Ptr a -> Ptr a -> CSize -> IO CInt
foreign import ccall safe   "sdr_memcpy_d2h"           c_d2h        :: Ptr Ctx -> Ptr a -> Ptr a -> CSize -> IO CInt
foreign import ccall safe   "sdr_memcpy_d2d"           c_d2d        :: Ptr Ctx -> Ptr a -> Ptr a -> CSize -> IO CInt
foreign import ccall safe   "sdr_decimator_create"     c_decCreate  :: Ptr Ctx -> CInt -> CInt -> Ptr CFloat -> CInt -> CInt -> Ptr (Ptr DecimatorH) -> IO CInt
foreign import ccall unsafe "sdr_decimator_num_coeffs" c_decNum     :: Ptr DecimatorH -> IO CInt
foreign import ccall safe   "sdr_decimate_one"         c_decOne     :: Ptr DecimatorH -> CInt -> Ptr a -> Ptr a -> CInt -> IO CInt
foreign import ccall safe   "sdr_decimate_cross"       c_decCross   :: Ptr DecimatorH -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> Ptr a -> CInt -> IO CInt
foreign import ccall unsafe "&sdr_decimator_destroy"   p_decDestroy :: FunPtr (Ptr DecimatorH -> IO ())

sdrDevice :: CInt
sdrDevice = 1

check :: IO CInt -> IO ()
check act = do
    st <- act
    unless (st == 0) $ c_lastError >>= peekCString >>= error

-- | One allocation (freed by the finalizer of its ForeignPtr-like box), an element offset and a length: slices share
--   the allocation, like storable vectors share their ForeignPtr.
data Block = Block { blockCtx :: Ptr Ctx, blockBase :: ForeignPtr () }

-- | immutable device vector
data DVector a    = DVector  !Block !Int !Int        -- block, offset (elements), length
-- | mutable device vector
data DMVector s a = DMVector !Block !Int !Int

type instance VG.Mutable DVector = DMVector

elemPtr :: forall a. Storable a => Block -> Int -> Ptr a
elemPtr b off = castPtr (unsafeForeignPtrToPtr (blockBase b)) `plusPtr` (off * sizeOf (undefined :: a))

newBlock :: Ptr Ctx -> Int -> IO Block
newBlock ctx bytes = do
    p  <- alloca $ \pp -> check (c_devAlloc ctx (fromIntegral (max bytes 16)) pp) >> peek pp
    fp <- FC.newForeignPtr p (c_devFree ctx p >> return ())   -- Haskell finalizer: it closes over the context
    return (Block ctx fp)

-- | one DMA of n elements starting at element `off` of the block into a fresh storable vector
download :: forall a. Storable a => Block -> Int -> Int -> IO (VS.Vector a)
download b off n = do
    mv <- VSM.new n
    VSM.unsafeWith mv $ \h -> do
        check (c_d2h (blockCtx b) h (elemPtr b off :: Ptr a) (fromIntegral (n * sizeOf (undefined :: a))))
        check (c_ctxSync (blockCtx b))
    VS.unsafeFreeze mv

instance Storable a => VGM.MVector DMVector a where
    basicLength (DMVector _ _ n)            = n
    basicUnsafeSlice i m (DMVector b o _)   = DMVector b (o + i) m          -- VGM.unsafeDrop (Filter.hs:547): O(1)
    basicOverlaps (DMVector b1 o1 n1) (DMVector b2 o2 n2) =
        blockBase b1 == blockBase b2 && o1 < o2 + n2 && o2 < o1 + n1
    basicUnsafeNew n                        = unsafePrimToPrim $ do         -- VGM.new (Filter.hs:513)
        ctx <- defaultCtx
        b   <- newBlock ctx (n * sizeOf (undefined :: a))
        return (DMVector b 0 n)
    basicInitialize _                       = return ()
    basicUnsafeRead (DMVector b o _) i      = unsafePrimToPrim $ alloca $ \h -> do
        check (c_d2h (blockCtx b) h (elemPtr b (o + i) :: Ptr a) (fromIntegral (sizeOf (undefined :: a))))
        check (c_ctxSync (blockCtx b))
        peek h
    basicUnsafeWrite (DMVector b o _) i x   = unsafePrimToPrim $ alloca $ \h -> do
        poke h x
        check (c_h2d (blockCtx b) (elemPtr b (o + i) :: Ptr a) h (fromIntegral (sizeOf (undefined :: a))))
        check (c_ctxSync (blockCtx b))
    basicUnsafeCopy (DMVector bd od n) (DMVector bs os _) = unsafePrimToPrim $
        check (c_d2d (blockCtx bd) (elemPtr bd od :: Ptr a) (elemPtr bs os :: Ptr a) (fromIntegral (n * sizeOf (undefined :: a))))

instance Storable a => VG.Vector DVector a where
    basicUnsafeFreeze (DMVector b o n)      = return (DVector b o n)        -- VG.unsafeFreeze (Filter.hs:519): no copy
    basicUnsafeThaw   (DVector b o n)       = return (DMVector b o n)
    basicLength (DVector _ _ n)             = n
    basicUnsafeSlice i m (DVector b o _)    = DVector b (o + i) m           -- VG.drop (Filter.hs:550, 592): O(1)
    basicUnsafeIndexM (DVector b o _) i     = return $! VS.head (unsafePerformIO (download b (o + i) 1))
    basicUnsafeCopy mv v                    = VG.basicUnsafeThaw v >>= VGM.basicUnsafeCopy mv

-- | upload a host vector (one DMA) -- the producer end of a device-resident pipeline
fromStorable :: forall a. Storable a => VS.Vector a -> IO (DVector a)
fromStorable v = do
    ctx <- defaultCtx
    b   <- newBlock ctx (VS.length v * sizeOf (undefined :: a))
    VS.unsafeWith v $ \h -> check (c_h2d ctx (elemPtr b 0 :: Ptr a) h (fromIntegral (VS.length v * sizeOf (undefined :: a))))
    check (c_ctxSync ctx)
    return (DVector b 0 (VS.length v))

-- | download (one DMA) -- the consumer end
toStorable :: Storable a => DVector a -> IO (VS.Vector a)
toStorable (DVector b o n) = download b o n

-- | one context (device 0, one stream) for every vector of the process: created on first use
theCtx :: Ptr Ctx
theCtx = unsafePerformIO $ alloca $ \pp -> check (c_ctxCreate 0 pp) >> peek pp
{-# NOINLINE theCtx #-}

defaultCtx :: IO (Ptr Ctx)
defaultCtx = return theCtx

-- | 'SDR.Filter.fastDecimatorC' (Filter.hs:352-356) over device vectors: the SAME record type, instantiated at
--   DVector, so @firDecimator deci 8192@ is the unmodified reference pipe moving HBM-resident buffers.  Both closures
--   are enqueue-only (layer 2 with SDR_DEVICE pointers); nothing crosses PCIe.
deviceDecimatorC :: Int -> [Float] -> IO (Decimator IO DVector DMVector (Complex Float))
deviceDecimatorC factor coeffs = do
    ctx <- defaultCtx
    h   <- alloca $ \pp -> do
        VS.unsafeWith (VS.fromList coeffs) $ \pc ->
            check (c_decCreate ctx 1 (fromIntegral factor) (castPtr pc) (fromIntegral (length coeffs)) 4 pp)
        peek pp
    fp  <- newForeignPtr p_decDestroy h
    n   <- fromIntegral <$> c_decNum h
    let one num (DVector bi oi _) (DMVector bo oo _) = withForeignPtr fp $ \hp ->
            check $ c_decOne hp (fromIntegral num) (elemPtr bi oi :: Ptr (Complex Float)) (elemPtr bo oo) sdrDevice
        cross num (DVector bl ol nl) (DVector bn on nn) (DMVector bo oo _) = withForeignPtr fp $ \hp ->
            check $ c_decCross hp (fromIntegral num) (elemPtr bl ol :: Ptr (Complex Float)) (fromIntegral nl)
                               (elemPtr bn on) (fromIntegral nn) (elemPtr bo oo) sdrDevice
    return $ Decimator n factor one cross
