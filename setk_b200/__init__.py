"""
setk_b200 -- B200-native (sm_100a) mask-based adaptive beamformer hot path of
funcwj/setk: multichannel STFT -> mask-weighted spatial covariance ->
MVDR / MPDR / GEV / PMWF weights -> beamform apply -> iSTFT, as hand-written
CUDA kernels behind the C-ABI in include/setk_b200.h.

    setk_b200.libs.{utils,stft,beamformer,data_handler,opts}   drop-in mirror of
                                                 scripts/sptk/libs of the reference
    setk_b200.plan      StftPlan + plan-free kernels on torch tensors
    setk_b200.engine    batched utterance pipeline (the benchmarked hot path)

No CPU fallback: importing is cheap, but the first kernel call loads
setk_b200/libsetk_b200.so and raises ImportError if it has not been built.
"""
__version__ = "0.1.0"
