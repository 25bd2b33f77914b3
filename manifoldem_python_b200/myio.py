"""Pickle in/out with the reference's semantics (modules/myio.py:21-59): fin1 returns the dict or
None on ANY failure; fout1 writes {key: value} with pickle.HIGHEST_PROTOCOL."""
import pickle


def fin1(filename):
    try:
        with open(filename, 'rb') as f:
            return pickle.load(f)
    except Exception:
        return None


def fout1(filename, key_list, v_list):
    with open(filename, 'wb') as f:
        pickle.dump(dict(zip(key_list, v_list)), f, protocol=pickle.HIGHEST_PROTOCOL)


def fout2(filename, d):
    fout1(filename, list(d.keys()), list(d.values()))
