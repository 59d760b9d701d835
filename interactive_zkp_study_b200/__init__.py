"""interactive_zkp_study_b200 -- B200-native (sm_100a) prover hot path for
tokamak-network/interactive-zkp-study: BN254 G1/G2 MSM and Fr NTT behind the reference's own
Python prover functions.  The compute lives in libzkp_b200.so (hand-written CUDA,
C ABI in include/zkp_b200.h); there is no CPU fallback."""

__all__ = ["native", "build"]
