"""mofo_b200 — B200-native (sm_100a) implementation of the MOFO masked-video-autoencoder pretraining step.

Public, reference-shaped surface:
  mofo_b200.masking_generator.TubeMaskingGenerator_BB / TubeMaskingGenerator
  mofo_b200.modeling_pretrain.pretrain_{mae_small,videomae_base,videomae_large}_patch16_224
  mofo_b200.engine_for_pretraining.train_one_epoch_BB
All device work runs in libmofo_sm100.so (hand-written CUDA, C ABI in include/mofo_b200.h).
"""
__version__ = "0.1.0"
