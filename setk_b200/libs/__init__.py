"""Drop-in mirror of the reference's scripts/sptk/libs for the beamformer hot path."""
