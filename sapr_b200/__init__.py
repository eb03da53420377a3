"""sapr_b200 -- B200-native (sm_100a) replacement for the HMM hot path of frankcholula/sapr assignment2.

Drop-in modules (same names and surfaces as the reference's): ``custom_hmm`` (class HMM),
``hmmlearn_hmm`` (GaussianHMM-style class + HMMLearnModel), ``mfcc_extract`` and ``decoder``.
Batched / multi-GPU surface: ``engine`` (PackedBatch, WordModels, train_words) and ``dist``.
All arithmetic runs in libsaprb200.so (csrc/*.cu) through the C ABI of include/sapr_b200.h; importing
this package does not need a GPU, calling it does (there is no CPU fallback).
"""
__version__ = "0.1.0"
