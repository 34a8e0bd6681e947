{-# LANGUAGE ForeignFunctionInterface, RecordWildCards #-}

-- | RMDF.FFI — `foreign import ccall` bindings of include/rmdf.h (librmdf_b200.so) for the
--   reference's own host code (ShaderRendering.hs / FrameBuffer.hs / HDREnvMap.hs).
--
--   NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC (SURVEY.md D8).  The
--   same exported symbols are exercised by the Python ctypes harness (tests/) and the C++
--   host (host/).  Link with:  ghc ... -lrmdf_b200 -L<dir>   (or `extra-libraries: rmdf_b200`
--   in rmdf.cabal).
--
--   Every call is `safe` (it may block on the GPU and may be issued from any OS thread under
--   -threaded, e.g. the mapConcurrently at ShaderRendering.hs:142); every entry point of the
--   library selects its own device and serialises on its context.
module RMDF.FFI
    ( RmdfCtx
    , FragmentShader(..)
    , withRmdf
    , rmdfRenderTile
    , rmdfConvolve
    , rmdfResize
    , rmdfBindLatLong
    , rmdfJulia
    ) where

import Control.Concurrent (runInBoundThread)
import Control.Exception (bracket)
import Control.Monad (when)
import Data.Int
import Data.Word
import Foreign.C.String
import Foreign.C.Types
import Foreign.Marshal.Alloc
import Foreign.Marshal.Array
import Foreign.Ptr
import Foreign.Storable
import qualified Data.Vector.Storable as VS
import qualified Data.Vector.Storable.Mutable as VSM

data RmdfCtxOpaque
type RmdfCtx = Ptr RmdfCtxOpaque

-- | Same constructor order as ShaderRendering.hs:46-47, so `fromEnum` is the C scene id.
data FragmentShader = FSDECornellBoxShader | FSDETestShader | FSMBPower8Shader | FSMBGeneralShader
                      deriving (Show, Eq, Enum)

foreign import ccall safe "rmdf_create"            c_create     :: Ptr RmdfCtx -> CInt -> IO CInt
foreign import ccall safe "rmdf_destroy"           c_destroy    :: RmdfCtx -> IO ()
foreign import ccall safe "rmdf_last_error"        c_last_error :: IO CString
foreign import ccall safe "rmdf_ctx_last_error"    c_ctx_last_error :: RmdfCtx -> CString -> CInt -> IO CInt
foreign import ccall safe "rmdf_render"            c_render     :: RmdfCtx -> Ptr RmdfParams -> Ptr Word32 -> IO CInt
foreign import ccall safe "rmdf_envmap_convolve"   c_convolve   :: RmdfCtx -> Ptr CFloat -> CInt -> CInt -> Ptr CFloat -> CInt -> Ptr CFloat -> IO CInt
foreign import ccall safe "rmdf_envmap_resize"     c_resize     :: RmdfCtx -> Ptr CFloat -> CInt -> CInt -> CInt -> Ptr CFloat -> IO CInt
foreign import ccall safe "rmdf_envmap_resize_height" c_resize_h :: CInt -> CInt -> CInt -> IO CInt
foreign import ccall safe "rmdf_envmap_bind"       c_bind       :: RmdfCtx -> CInt -> Ptr CFloat -> CInt -> CInt -> IO CInt
foreign import ccall safe "rmdf_cornell_box_vertices" c_cornell :: Ptr CFloat -> IO CInt
foreign import ccall safe "rmdf_set_cornell"       c_set_cornell :: RmdfCtx -> Ptr CFloat -> CInt -> IO CInt
foreign import ccall safe "rmdf_julia"             c_julia      :: RmdfCtx -> CInt -> CInt -> CInt -> CDouble -> CInt -> CInt -> CInt -> Ptr Word32 -> IO CInt

-- | rmdf_params (include/rmdf.h): eighteen 32-bit fields.
data RmdfParams = RmdfParams
    { pScene, pFbW, pFbH :: !Int32
    , pTime              :: !Float
    , pTileIdx, pResolveLog2, pOutFormat, pMode, pLodPolicy, pShardRank, pShardCount, pLayout :: !Int32
    }

instance Storable RmdfParams where
    sizeOf    _ = 72
    alignment _ = 4
    peek _      = error "RmdfParams: write-only"
    poke p RmdfParams { .. } = do
        let i off v = pokeByteOff p off (v :: Int32)
        i 0 pScene >> i 4 pFbW >> i 8 pFbH >> pokeByteOff p 12 pTime
        i 16 pTileIdx >> i 20 pResolveLog2 >> i 24 pOutFormat >> i 28 pMode
        i 32 pLodPolicy >> i 36 pShardRank >> i 40 pShardCount >> i 44 pLayout
        mapM_ (\off -> i off 0) [48, 52, 56, 60, 64, 68]            -- shard_grid_x/y, shard_skew (0 = 8×8 grid), flags (0 = the reference's build), row_begin/row_end (0,0 = all rows)

-- | `Left <message>` like the ExceptT String the reference threads through withShaderRenderer.
--   The message comes through the context (rmdf_ctx_last_error): under -threaded an unbound Haskell thread may run
--   the failing `safe` call and the fetch on different OS threads, where the thread-local rmdf_last_error is stale.
checkCtx :: RmdfCtx -> CInt -> IO (Either String ())
checkCtx _   0 = return $ Right ()
checkCtx ctx _ = allocaBytes 512 $ \buf -> c_ctx_last_error ctx buf 512 >> (Left <$> peekCString buf)

-- | For the calls that have no context yet (rmdf_create, rmdf_cornell_box_vertices): call and fetch on ONE OS thread.
check :: CInt -> IO (Either String ())
check 0 = return $ Right ()
check _ = Left <$> (c_last_error >>= peekCString)

orThrowBound :: IO CInt -> IO ()
orThrowBound act = runInBoundThread (act >>= check) >>= either (ioError . userError) return

orThrow :: IO CInt -> IO ()
orThrow act = act >>= check >>= either (ioError . userError) return

-- | Bracket-style lifetime, the shape of withShaderRenderer (ShaderRendering.hs:60).  Uploads the
--   Cornell geometry (mkCornellBoxVerticesTex, CornellBox.hs:21-46).
withRmdf :: Int -> (RmdfCtx -> IO a) -> IO a
withRmdf device = bracket create c_destroy
  where create = alloca $ \pp -> do
            orThrowBound $ c_create pp (fromIntegral device)
            ctx <- peek pp
            allocaArray 288 $ \xyz -> do
                orThrowBound $ c_cornell xyz
                c_set_cornell ctx xyz 96 >>= checkCtx ctx >>= either (ioError . userError) return
            return ctx

-- | The body of drawShaderTile (ShaderRendering.hs:151-196) for a frame buffer handed out by
--   fillFrameBuffer (FrameBuffer.hs:117-158): renders tile `tileIdx` (Nothing = whole frame) of a
--   w×h buffer into the mapped PBO vector.
rmdfRenderTile :: RmdfCtx -> FragmentShader -> Maybe Int -> Int -> Int -> Double -> VSM.IOVector Word32 -> IO (Either String ())
rmdfRenderTile ctx shd tileIdx w h time fb
    | VSM.length fb < w * h = return . Left $ "rmdfRenderTile: frame buffer too small"
    | otherwise =
        alloca $ \pp -> do
            poke pp RmdfParams { pScene = fromIntegral $ fromEnum shd
                               , pFbW = fromIntegral w, pFbH = fromIntegral h
                               , pTime = realToFrac time                -- ShaderRendering.hs:171
                               , pTileIdx = maybe (-1) fromIntegral tileIdx
                               , pResolveLog2 = 0, pOutFormat = 0, pMode = 0, pLodPolicy = 0
                               , pShardRank = 0, pShardCount = 1, pLayout = 0 }
            VSM.unsafeWith fb $ \p -> c_render ctx pp p >>= checkCtx ctx

-- | cosineConvolveHDREnvMap (HDREnvMap.hs:217-254) for all powers at once; images are the
--   interleaved-RGB `VS.Vector Float` of JuicyPixels' `Image PixelRGBF` (imageData).
rmdfConvolve :: RmdfCtx -> VS.Vector Float -> Int -> Int -> [Float] -> IO (Either String [VS.Vector Float])
rmdfConvolve ctx src w h powers = do
    let n = w * h * 3
    out <- VSM.new (n * length powers)
    r <- VS.unsafeWith src $ \ps -> withArrayLen (map realToFrac powers) $ \np pp ->
           VSM.unsafeWith out $ \po ->
             c_convolve ctx (castPtr ps) (fromIntegral w) (fromIntegral h) pp (fromIntegral np) (castPtr po) >>= checkCtx ctx
    frozen <- VS.unsafeFreeze out
    return $ fmap (const [VS.slice (i * n) n frozen | i <- [0 .. length powers - 1]]) r

-- | resizeHDRImage (HDREnvMap.hs:169-195); returns (height, pixels).
rmdfResize :: RmdfCtx -> VS.Vector Float -> Int -> Int -> Int -> IO (Either String (Int, VS.Vector Float))
rmdfResize ctx src w h dstw = do
    dsth <- fromIntegral <$> c_resize_h (fromIntegral w) (fromIntegral h) (fromIntegral dstw)
    out  <- VSM.new (dstw * dsth * 3)
    r <- VS.unsafeWith src $ \ps -> VSM.unsafeWith out $ \po ->
           c_resize ctx (castPtr ps) (fromIntegral w) (fromIntegral h) (fromIntegral dstw) (castPtr po) >>= checkCtx ctx
    frozen <- VS.unsafeFreeze out
    return $ fmap (const (dsth, frozen)) r

-- | latLongHDREnvMapToCubeMap + setTextureShader (HDREnvMap.hs:118-163, ShaderRendering.hs:173-174):
--   slot 0..4 = env_reflection, env_cos_1, env_cos_8, env_cos_64, env_cos_512.
rmdfBindLatLong :: RmdfCtx -> Int -> VS.Vector Float -> Int -> Int -> IO (Either String ())
rmdfBindLatLong ctx slot img w h =
    VS.unsafeWith img $ \p -> c_bind ctx (fromIntegral slot) (castPtr p) (fromIntegral w) (fromIntegral h) >>= checkCtx ctx

-- | juliaAnimated (Fractal2D.hs:63-98) into a fillFrameBuffer vector.
rmdfJulia :: RmdfCtx -> Int -> Int -> VSM.IOVector Word32 -> Bool -> Double -> IO (Either String ())
rmdfJulia ctx w h fb smooth tick = do
    when (VSM.length fb < w * h) $ ioError (userError "rmdfJulia: frame buffer too small")
    VSM.unsafeWith fb $ \p ->
        c_julia ctx (fromIntegral w) (fromIntegral h) (if smooth then 1 else 0) (realToFrac tick)
                0 (fromIntegral h) 0 p >>= checkCtx ctx
