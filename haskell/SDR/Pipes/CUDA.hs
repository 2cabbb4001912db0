{-# LANGUAGE ForeignFunctionInterface, ScopedTypeVariables, ExistentialQuantification #-}
{-| Device-resident members of the streaming stages of "SDR.Filter", "SDR.Demod", "SDR.Util" and the producer /
    consumer edges of "SDR.Serialize" / "SDR.NetworkStream", over layer 3 of libsdr_b200's C ABI
    (include/sdr_b200.h: @sdr_pipe_*@).  Each stage is an ordinary @Pipe (VS.Vector a) (VS.Vector b) IO ()@, so it
    composes with the unchanged Pipes glue ('>->', "SDR.PipeUtils"); stages joined with 'connect' additionally hand
    their vectors to each other inside HBM and only the ends of the chain touch host memory.

    Source only: GHC is not in the build image, so this module has never been compiled.  It is kept mechanical --
    one @foreign import@ per C entry point -- and mirrors what the Python host mirror (sdr_b200/filter.py,
    sdr_b200/util.py, sdr_b200/serialize.py) does and the tests exercise.

    Reference counterparts: firFilter / firDecimator / firResampler hs_sources/SDR/Filter.hs:532-727, fmDemod
    hs_sources/SDR/Demod.hs:38-46, dcBlockingFilter Filter.hs:730-739, interleavedIQUnsignedByteToFloat
    hs_sources/SDR/Util.hs:104, fromHandle / toHandle hs_sources/SDR/Serialize.hs:78-83, udpSource / udpSink
    hs_sources/SDR/NetworkStream.hs:28-42.
-}
module SDR.Pipes.CUDA (
    Stage,
    -- * Stages
    cudaFirFilter, cudaFirDecimator, cudaFirResampler, cudaFmDemod, cudaFmFrontEnd, cudaU8Decimator, cudaFmLowRate,
    cudaConvertU8, cudaScale, cudaDcBlockingFilter,
    -- * Composition
    connect, setBatch, stagePipe,
    -- * Checkpoint / resume
    stateSave, stateRestore,
    -- * Whole chains on file descriptors
    runHandles, runUdp
    ) where

import           Control.Monad                (forever, unless, when)
import qualified Data.ByteString              as BS
import qualified Data.ByteString.Unsafe       as BS
import           Data.IORef
import qualified Data.Vector.Storable         as VS
import qualified Data.Vector.Storable.Mutable as VSM
import           Foreign.C.String
import           Foreign.C.Types
import           Foreign.ForeignPtr
import           Foreign.Marshal.Alloc
import           Foreign.Ptr
import           Foreign.Storable
import           Pipes
import           System.Posix.Types           (Fd (..))

import           SDR.Filter.CUDA              (Ctx, DecimatorH, FilterH, ResamplerH)

data PipeH

-- | A stage handle.  'stageKeep' holds everything the C side refers to by raw pointer and that must therefore outlive
--   the stage: the plugin record it was built from (the library keeps @FirRec *@ / @ResRec *@ inside the pipe) and every
--   stage connected downstream ('connect': the library keeps a raw @downstream@ pointer and forwards into it).  Each
--   use of the stage touches the list after the foreign call, so the GC cannot finalize a kept object while a call
--   that may reach it is running.  (sdr_pipe_destroy also unlinks a stage from its neighbours, so finalization ORDER
--   at program exit is harmless.)
data Stage a b = Stage { stageH :: ForeignPtr PipeH, stageKeep :: IORef [Keep] }
data Keep = forall x. Keep (ForeignPtr x)

touchStage :: Stage a b -> IO ()
touchStage st = readIORef (stageKeep st) >>= mapM_ (\(Keep fp) -> touchForeignPtr fp) >> touchForeignPtr (stageH st)

-- | run a foreign call on the stage's handle, keeping everything it depends on alive across the call
withStage :: Stage a b -> (Ptr PipeH -> IO r) -> IO r
withStage st act = do
    r <- withForeignPtr (stageH st) act
    touchStage st
    return r

foreign import ccall unsafe "sdr_last_error"          c_lastError   :: IO CString
foreign import ccall safe   "sdr_pipe_fir_filter"     c_pipeFilter  :: Ptr FilterH -> CInt -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_fir_decimator"  c_pipeDecim   :: Ptr DecimatorH -> CInt -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_fir_resampler"  c_pipeResamp  :: Ptr ResamplerH -> CInt -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_fm_frontend"    c_pipeFmFront :: Ptr DecimatorH -> CInt -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_u8_decimator"   c_pipeU8Dec   :: Ptr DecimatorH -> CInt -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_fm_lowrate"     c_pipeFmLow   :: Ptr ResamplerH -> CInt -> Ptr FilterH -> CInt -> CFloat -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_state_size"     c_stateSize   :: Ptr PipeH -> Ptr CSize -> IO CInt
foreign import ccall safe   "sdr_pipe_state_save"     c_stateSave   :: Ptr PipeH -> Ptr a -> CSize -> Ptr CSize -> IO CInt
foreign import ccall safe   "sdr_pipe_state_restore"  c_stateRestore :: Ptr PipeH -> Ptr a -> CSize -> IO CInt
foreign import ccall safe   "sdr_pipe_fm_demod"       c_pipeFmDemod :: Ptr Ctx -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_convert_u8"     c_pipeConvert :: Ptr Ctx -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_scale"          c_pipeScale   :: Ptr Ctx -> CFloat -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall safe   "sdr_pipe_dc_blocker"     c_pipeDc      :: Ptr Ctx -> Ptr (Ptr PipeH) -> IO CInt
foreign import ccall unsafe "&sdr_pipe_destroy"       p_pipeDestroy :: FunPtr (Ptr PipeH -> IO ())
foreign import ccall safe   "sdr_pipe_push"           c_push        :: Ptr PipeH -> Ptr a -> CLLong -> CInt -> IO CInt
foreign import ccall unsafe "sdr_pipe_ready"          c_ready       :: Ptr PipeH -> Ptr CInt -> IO CInt
foreign import ccall unsafe "sdr_pipe_next_len"       c_nextLen     :: Ptr PipeH -> Ptr CLLong -> IO CInt
foreign import ccall safe   "sdr_pipe_pop"            c_pop         :: Ptr PipeH -> Ptr b -> Ptr CLLong -> CInt -> IO CInt
foreign import ccall unsafe "sdr_pipe_connect"        c_connect     :: Ptr PipeH -> Ptr PipeH -> IO CInt
foreign import ccall safe   "sdr_pipe_set_batch"      c_setBatch    :: Ptr PipeH -> CLLong -> IO CInt
-- sdr_io_stats_t is six 8-byte fields; the binding only needs the storage
foreign import ccall safe   "sdr_pipe_run_fd"         c_runFd       :: Ptr PipeH -> Ptr PipeH -> CInt -> CLLong -> CLLong -> CInt -> CInt -> Ptr () -> IO CInt

sdrHost :: CInt
sdrHost = 0

check :: IO CInt -> IO ()
check act = do
    st <- act
    unless (st == 0) $ c_lastError >>= peekCString >>= error

mk :: [Keep] -> (Ptr (Ptr PipeH) -> IO CInt) -> IO (Stage a b)
mk keep create = do
    h  <- alloca $ \pp -> check (create pp) >> peek pp
    fp <- newForeignPtr p_pipeDestroy h
    Stage fp <$> newIORef keep

-- | 'SDR.Filter.firFilter' (Filter.hs:532): the record handles come from "SDR.Filter.CUDA" as ForeignPtrs and are kept
--   alive by the stage (the library refers to the record for as long as the stage exists)
cudaFirFilter :: ForeignPtr FilterH -> Int -> IO (Stage a a)
cudaFirFilter f blockSizeOut = withForeignPtr f $ \fp -> mk [Keep f] (c_pipeFilter fp (fromIntegral blockSizeOut))

-- | 'SDR.Filter.firDecimator' (Filter.hs:574)
cudaFirDecimator :: ForeignPtr DecimatorH -> Int -> IO (Stage a a)
cudaFirDecimator d blockSizeOut = withForeignPtr d $ \dp -> mk [Keep d] (c_pipeDecim dp (fromIntegral blockSizeOut))

-- | 'SDR.Filter.firResampler' (Filter.hs:679)
cudaFirResampler :: ForeignPtr ResamplerH -> Int -> IO (Stage a a)
cudaFirResampler r blockSizeOut = withForeignPtr r $ \rp -> mk [Keep r] (c_pipeResamp rp (fromIntegral blockSizeOut))

-- | 'SDR.Demod.fmDemod' (Demod.hs:40)
cudaFmDemod :: Ptr Ctx -> IO (Stage (Complex' Float) Float)
cudaFmDemod ctx = mk [] (c_pipeFmDemod ctx)

-- | @P.map interleavedIQUnsignedByteToFloat >-> firDecimator d n >-> fmDemod@ (examples/fm/fm.hs:34-37) as one kernel
cudaFmFrontEnd :: ForeignPtr DecimatorH -> Int -> IO (Stage CUChar Float)
cudaFmFrontEnd d blockSizeOut = withForeignPtr d $ \dp -> mk [Keep d] (c_pipeFmFront dp (fromIntegral blockSizeOut))

-- | @P.map interleavedIQUnsignedByteToFloat >-> firDecimator d n@ (fm.hs:34-36) as one kernel: u8 IQ in, complex out
cudaU8Decimator :: ForeignPtr DecimatorH -> Int -> IO (Stage CUChar (Complex' Float))
cudaU8Decimator d blockSizeOut = withForeignPtr d $ \dp -> mk [Keep d] (c_pipeU8Dec dp (fromIntegral blockSizeOut))

-- | @firResampler r nr >-> firFilter f n >-> P.map (VG.map (* k))@ (fm.hs:38-40) as one kernel
cudaFmLowRate :: ForeignPtr ResamplerH -> Int -> ForeignPtr FilterH -> Int -> Float -> IO (Stage Float Float)
cudaFmLowRate r blockResampler f blockSizeOut k = withForeignPtr r $ \rp -> withForeignPtr f $ \fp ->
    mk [Keep r, Keep f] (c_pipeFmLow rp (fromIntegral blockResampler) fp (fromIntegral blockSizeOut) (realToFrac k))

-- | @P.map interleavedIQUnsignedByteToFloat@ (Util.hs:104)
cudaConvertU8 :: Ptr Ctx -> IO (Stage CUChar (Complex' Float))
cudaConvertU8 ctx = mk [] (c_pipeConvert ctx)

-- | @P.map (VG.map (* k))@ (fm.hs:40)
cudaScale :: Ptr Ctx -> Float -> IO (Stage Float Float)
cudaScale ctx k = mk [] (c_pipeScale ctx (realToFrac k))

-- | 'SDR.Filter.dcBlockingFilter' (Filter.hs:730)
cudaDcBlockingFilter :: Ptr Ctx -> IO (Stage Float Float)
cudaDcBlockingFilter ctx = mk [] (c_pipeDc ctx)

-- the binding's stand-in for Data.Complex.Complex so this file needs no extra imports to read
type Complex' a = (a, a)

-- | '>->' on the device: the vectors @src@ yields are awaited by @dst@ without leaving HBM.  @src@ keeps @dst@ (and,
--   transitively, everything @dst@ keeps) alive: the library forwards into @dst@ through a raw pointer.
connect :: Stage a b -> Stage b c -> IO ()
connect s d = do
    withStage s $ \sp -> withStage d $ \dp -> check (c_connect sp dp)
    dk <- readIORef (stageKeep d)
    modifyIORef' (stageKeep s) ((Keep (stageH d) : dk) ++)

-- | launch only once this many outputs are computable (latency for throughput; yielded vectors are unchanged)
setBatch :: Stage a b -> Int -> IO ()
setBatch s n = withStage s $ \sp -> check (c_setBatch sp (fromIntegral n))

-- | the stage's carried stream state (tail, counters, carried samples, un-popped outputs) as a blob: sdr_pipe_state_save
stateSave :: Stage a b -> IO BS.ByteString
stateSave s = withStage s $ \sp -> do
    n <- alloca $ \pn -> check (c_stateSize sp pn) >> peek pn
    allocaBytes (fromIntegral n) $ \buf -> alloca $ \pw -> do
        check (c_stateSave sp buf n pw)
        w <- peek pw
        BS.packCStringLen (castPtr buf, fromIntegral w)

-- | load a saved state into a stage constructed the same way: the stream continues bit for bit (sdr_pipe_state_restore)
stateRestore :: Stage a b -> BS.ByteString -> IO ()
stateRestore s blob = withStage s $ \sp -> BS.unsafeUseAsCStringLen blob $ \(p, n) -> check (c_stateRestore sp p (fromIntegral n))

-- | A chain of connected stages as one Pipe: vectors awaited here are pushed into @headS@, every vector @sinkS@ has
--   ready is yielded (for a single stage pass it twice).  Host vectors in, host vectors out, like the reference.
stagePipe :: forall a b c d. (Storable a, Storable d) => Stage a b -> Stage c d -> Pipe (VS.Vector a) (VS.Vector d) IO ()
stagePipe headS sinkS = forever $ do
    v <- await
    lift $ withStage headS $ \hp -> VS.unsafeWith v $ \p ->
        check (c_push hp p (fromIntegral (VS.length v)) sdrHost)
    drain
  where
    drain = do
        n <- lift $ withStage sinkS $ \sp -> alloca $ \pn -> check (c_ready sp pn) >> peek pn
        when (n > 0) $ do
            out <- lift $ withStage sinkS $ \sp -> do
                len <- alloca $ \pl -> check (c_nextLen sp pl) >> peek pl
                buf <- VSM.new (fromIntegral len)
                VSM.unsafeWith buf $ \op -> alloca $ \pl -> check (c_pop sp op pl sdrHost)
                VS.unsafeFreeze buf
            yield out
            drain

-- | @runEffect $ fromHandle samples hIn >-> chain >-> toHandle hOut@ (Serialize.hs:78-83) as one native loop:
--   read() lands in a page-locked ring, nothing crosses the Haskell heap.  @Nothing@ discards the output.
runHandles :: Stage a b -> Stage c d -> Int -> Fd -> Maybe Fd -> IO ()
runHandles h s samples (Fd fin) fout =
    withStage h $ \hp -> withStage s $ \sp -> allocaBytes 48 $ \st ->
        check (c_runFd hp sp fin (fromIntegral samples) 0 (maybe (-1) (\(Fd o) -> o) fout) 0 st)

-- | @runEffect $ udpSource sock size >-> chain >-> udpSink@ (NetworkStream.hs:28-42) for @count@ datagrams; the sink
--   descriptor must be a connected datagram socket
runUdp :: Stage a b -> Stage c d -> Fd -> Int -> Int -> Maybe Fd -> IO ()
runUdp h s (Fd sock) elemsPerDatagram count fout =
    withStage h $ \hp -> withStage s $ \sp -> allocaBytes 48 $ \st ->
        check (c_runFd hp sp sock (fromIntegral elemsPerDatagram) (fromIntegral count) (maybe (-1) (\(Fd o) -> o) fout)
                       (maybe 1 (const 3) fout) st)
