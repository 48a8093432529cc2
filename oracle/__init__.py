"""
oracle/ -- CPU restatement of the reference's mask-beamformer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (setk_b200/) may
import this directory; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and there only as the checker or
as the CPU arm that is timed beside the GPU number.

Parity status: PINNED.  The restatement is checked (oracle/make_golden.py,
tests/test_oracle_golden.py) against
  * the reference's own modules (scripts/sptk/libs/{utils,beamformer}.py)
    imported from /root/reference under the compatibility shim in
    oracle/ref_shim.py, on the doc example and on seeded synthetic input;
  * the reference's shipped end-to-end vectors
    doc/adaptive_beamformer/asset/{egs -> pmwf-0, pmwf-0-eig, pmwf-0-gev,
    gevd, mvdr}.wav  (PMWF chain: <= 1 LSB of PCM-16).
The STFT arithmetic itself lives in librosa==0.8.1 (requirements.txt:2),
which is not vendored in the reference and not installable here; its
published algorithm is restated in oracle/stft_oracle.py and pinned by the
+-1 LSB replay of the doc vectors above.
"""
