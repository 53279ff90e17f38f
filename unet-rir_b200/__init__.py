"""unet-rir on B200: the U-Net amp/phase hot path of igmsalinas/unet-rir as hand-written sm_100a
CUDA behind a C-ABI (include/urir.h), with the reference's Python surface on top.

Modules mirror the reference's flat file names: dl_models.u_net (UNet), amp_phase_trainer (Trainer),
preprocess, postprocess, datageneratorv2, dataset, rir_generation, main_training.
"""
__version__ = "0.1.0"
