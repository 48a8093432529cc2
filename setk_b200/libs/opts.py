"""
setk_b200.libs.opts -- argparse pieces shared by the CLIs; same flags and
defaults as the reference's scripts/sptk/libs/opts.py:9-49.
"""
import argparse


def strtobool(val):
    """distutils.util.strtobool (gone from the stdlib in 3.12): 1/0 for the usual spellings."""
    val = str(val).lower()
    if val in ("y", "yes", "t", "true", "on", "1"):
        return 1
    if val in ("n", "no", "f", "false", "off", "0"):
        return 0
    raise ValueError(f"invalid truth value {val!r}")


def str2tuple(string, sep=","):
    """Map "1.0,2,0" => (1.0, 2.0, 0.0)  (opts.py:9-18)"""
    return tuple(map(float, string.split(sep)))


class StftParser(object):
    """STFT argparser (opts.py:21-49)"""
    parser = argparse.ArgumentParser(add_help=False)
    parser.add_argument("--frame-len", type=int, default=512,
                        help="Frame length in number of samples (related to sample frequency)")
    parser.add_argument("--frame-hop", type=int, default=256,
                        help="Frame shift in number of samples (related to sample frequency)")
    parser.add_argument("--center", type=strtobool, default=True,
                        help="Value of parameter 'center' in librosa.stft functions")
    parser.add_argument("--round-power-of-two", type=strtobool, default=True,
                        help="If true, pad fft size to power of two")
    parser.add_argument("--window", type=str, default="hann",
                        help="Type of window function, see scipy.signal.get_window")
