"""Readers of the exported feature / match files -- same functions as imcui/hloc/utils/io.py:24-78 over utils/store.py
(HDF5 through h5py when installed, the built-in record file otherwise)."""
import numpy as np

from .parsers import names_to_pair, names_to_pair_old
from .store import open_store


def list_h5_names(path):
    """names of all groups that hold datasets (io.py:24-34)"""
    with open_store(path, "r") as st:
        return st.groups()


def get_keypoints(path, name, return_uncertainty=False):
    """io.py:37-47"""
    with open_store(path, "r") as st:
        p = st.read(name, "keypoints")
        unc = st.attrs(name, "keypoints").get("uncertainty")
    return (p, unc) if return_uncertainty else p


def find_pair(store, name0, name1):
    """key under which a pair was stored and whether it is reversed (io.py:50-67): new '/' keys first, then the old '_' ones"""
    for key_fn in (names_to_pair, names_to_pair_old):
        for rev, (a, b) in enumerate(((name0, name1), (name1, name0))):
            key = key_fn(a, b)
            if key in store:
                return key, bool(rev)
    raise ValueError(f"Could not find pair {(name0, name1)}... Maybe you matched with a different list of pairs? ")


def get_matches(path, name0, name1):
    """[K,2] index pairs and their scores (io.py:70-82); indices are flipped when the pair was stored reversed"""
    with open_store(path, "r") as st:
        pair, reverse = find_pair(st, name0, name1)
        m = st.read(pair, "matches0")
        sc = st.read(pair, "matching_scores0")
    idx = np.where(m != -1)[0]
    matches = np.stack([idx, m[idx]], -1)
    if reverse:
        matches = np.flip(matches, -1)
    return matches, sc[idx]
