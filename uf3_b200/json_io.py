"""JSON (de)serialisation of interaction-keyed maps.

File-format mirror of `/root/reference/uf3/util/json_io.py:11-83`: tuple keys
are written as dash-joined strings ("W-W-W"), arrays as lists; on load, lists
become arrays (a list of lists becomes a list of row arrays) and dash-joined
keys become tuples again.  Model files written by either implementation load
in the other.
"""
import json

import numpy as np


def encode_interaction_map(interaction_map):
    encoded = {}
    for key, value in interaction_map.items():
        if isinstance(value, dict):
            value = encode_interaction_map(value)
        elif isinstance(value, np.ndarray):
            value = value.tolist()
        elif isinstance(value, (list, tuple)):
            value = [v.tolist() if isinstance(v, np.ndarray) else v for v in value]
        elif isinstance(value, np.generic):
            value = value.item()
        if isinstance(key, tuple):
            key = "-".join(str(part) for part in key)
        encoded[key] = value
    return encoded


def decode_interaction_map(formatted_map):
    decoded = {}
    for key, value in formatted_map.items():
        if isinstance(value, dict):
            value = decode_interaction_map(value)
        elif isinstance(value, list):
            if len(value) > 0 and isinstance(value[0], list):
                value = [np.array(row) for row in value]
            else:
                value = np.array(value)
        if "-" in key:
            parts = key.split("-")
            try:
                parts = [int(p) for p in parts]
            except ValueError:
                pass
            key = tuple(parts)
        decoded[key] = value
    return decoded


def dump_interaction_map(interaction_map, indent=4, filename=None, write=False):
    text = json.dumps(encode_interaction_map(interaction_map), indent=indent)
    if write:
        with open(filename, "w") as handle:
            handle.write(text)
        return None
    return text


def load_interaction_map(filename):
    with open(filename, "r") as handle:
        return decode_interaction_map(json.load(handle))
