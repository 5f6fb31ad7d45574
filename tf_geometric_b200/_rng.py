# coding=utf-8
"""Seeds for the counter-based dropout / sampling kernels.

The reference relies on TensorFlow's and numpy's global generators (tf.nn.dropout in gcn.py:262 / gat.py:85,
tf.random.uniform and np.random.choice in utils/graph_utils.py:741-841).  Here every random operator call takes a
64-bit Philox key; callers may pass one explicitly (`seed=`) for reproducible runs, otherwise it is derived from a
process-wide base seed and a call counter, so consecutive calls draw independent masks."""
import threading

_lock = threading.Lock()
_state = {"seed": 0x5DEECE66D, "calls": 0}
_MASK64 = (1 << 64) - 1


def set_seed(seed):
    """Reset the process-wide base seed (the analogue of tf.random.set_seed / np.random.seed)."""
    with _lock:
        _state["seed"] = int(seed) & _MASK64
        _state["calls"] = 0


def next_seed():
    """A fresh 64-bit key: splitmix64 of (base seed + call counter)."""
    with _lock:
        _state["calls"] += 1
        z = (_state["seed"] + 0x9E3779B97F4A7C15 * _state["calls"]) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def resolve(seed):
    return next_seed() if seed is None else int(seed) & _MASK64
