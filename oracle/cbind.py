"""ctypes binding of oracle/libaz_oracle.so (CPU ORACLE -- test infrastructure only)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaz_oracle.so")

OZ_MAX_CELLS = 64
OZ_MAX_LEGAL = 64
GAME_CONNECT_FOUR = 0
GAME_BREAKTHROUGH = 1


class OzState(C.Structure):
    _fields_ = [
        ("game", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
        ("player", C.c_int32), ("ply", C.c_int32), ("winner", C.c_int32),
        ("pieces", C.c_int32 * 2),
        ("cell", C.c_int8 * OZ_MAX_CELLS),
    ]


class OzSelfplayCfg(C.Structure):
    _fields_ = [
        ("game", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
        ("n_playouts", C.c_int32),
        ("c_puct", C.c_double), ("dirichlet_ratio", C.c_double),
        ("use_dirichlet", C.c_int32), ("sample_moves", C.c_int32),
        ("num_probabilistic_actions", C.c_int32), ("keep_tree", C.c_int32),
        ("eval_kind", C.c_int32), ("eval_shift", C.c_int32),
        ("seed", C.c_uint64),
        ("start_random_plies_mod", C.c_int32), ("max_plies", C.c_int32),
    ]


class OzPlyRecord(C.Structure):
    _fields_ = [
        ("tree", C.c_int32), ("game_seq", C.c_int32), ("ply", C.c_int32), ("action", C.c_int32),
        ("n_legal", C.c_int32), ("player", C.c_int32),
        ("bb", C.c_uint64 * 2),
        ("root_q", C.c_double), ("v_a0c", C.c_double), ("v_offpolicy", C.c_double),
        ("root_n", C.c_int64),
        ("counts", C.c_int32 * OZ_MAX_LEGAL),
    ]


EVAL_FN = C.CFUNCTYPE(None, C.POINTER(OzState), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = os.path.join(_HERE, "az_oracle.c")
    hdr = os.path.join(_HERE, "az_oracle.h")
    if not force and os.path.exists(_LIB_PATH):
        newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
        if os.path.getmtime(_LIB_PATH) >= newest:
            return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libaz_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    P = C.POINTER
    sig = {
        "oz_init": (None, [P(OzState), C.c_int, C.c_int, C.c_int]),
        "oz_num_actions": (C.c_int, [P(OzState)]),
        "oz_legal": (C.c_int, [P(OzState), P(C.c_int32)]),
        "oz_apply": (C.c_int, [P(OzState), C.c_int]),
        "oz_terminal": (C.c_int, [P(OzState)]),
        "oz_current_player": (C.c_int, [P(OzState)]),
        "oz_returns": (None, [P(OzState), P(C.c_double)]),
        "oz_normalized_vector": (None, [P(OzState), P(C.c_float)]),
        "oz_board": (None, [P(OzState), P(C.c_double)]),
        "oz_bitboards": (None, [P(OzState), P(C.c_uint64)]),
        "oz_mix64": (C.c_uint64, [C.c_uint64]),
        "oz_counter": (C.c_uint64, [C.c_uint64] * 6),
        "oz_synth_eval": (None, [P(OzState), C.c_int, C.c_uint64, C.c_int, P(C.c_double), P(C.c_double)]),
        "oz_synth_eval_bb": (None, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p,
                                    C.c_void_p]),
        "oz_tree_new": (C.c_void_p, [C.c_int, C.c_double, C.c_int, C.c_int, C.c_double]),
        "oz_tree_free": (None, [C.c_void_p]),
        "oz_tree_reset": (None, [C.c_void_p]),
        "oz_tree_search": (None, [C.c_void_p, P(OzState), EVAL_FN, C.c_void_p, P(C.c_double), P(C.c_int64)]),
        "oz_tree_update_root": (None, [C.c_void_p, C.c_int]),
        "oz_tree_root_n": (C.c_int64, [C.c_void_p]),
        "oz_tree_root_q": (C.c_double, [C.c_void_p]),
        "oz_tree_root_children": (None, [C.c_void_p, P(C.c_int64), P(C.c_double), P(C.c_double)]),
        "oz_tree_value_a0c": (C.c_double, [C.c_void_p]),
        "oz_tree_value_offpolicy": (C.c_double, [C.c_void_p]),
        "oz_tree_counters": (None, [C.c_void_p, P(C.c_uint64)]),
        "oz_selfplay_game": (C.c_int, [P(OzSelfplayCfg), C.c_uint64, C.c_uint64, P(OzPlyRecord), C.c_int,
                                       P(C.c_double), P(C.c_uint64)]),
        "oz_start_position": (None, [P(OzSelfplayCfg), C.c_uint64, C.c_uint64, P(OzState)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
