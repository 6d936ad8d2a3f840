"""evcdiff -- B200-native (sm_100a) conditional video-diffusion sampling path.

Mirrors the reference's Python entry points for the hot path only:
  evcdiff.models            ddpm_sampler / ddim_sampler / FPNDM_sampler / get_sigmas   (models/__init__.py)
  evcdiff.models.pndm       runge_kutta / transfer / gen_order_1 / gen_order_4          (models/pndm.py)
  evcdiff.models.better.ncsnpp_more   UNetMore_DDPM / NCSNpp                            (models/better/ncsnpp_more.py)
  evcdiff.models.unet       UNet_DDPM / UNet                                            (models/unet.py)
All math runs in libevcdiff.so (hand-written CUDA, C ABI in include/evcdiff.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
