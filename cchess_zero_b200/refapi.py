"""The reference's `main.py` module namespace, served by this package.

`main.py` (chengstone/cchess-zero) keeps everything in one module: label tables and helpers (main.py:23-91, 208-232),
`leaf_node` / `MCTS_tree` (93-577), `GameBoard` (579-1109), `softmax` (1111-1116) and the caller, `class cchess_main`
(1118-1554), which only ever refers to those module-level names.  The drop-in contract (BASELINE north_star:
"main.py's train loop drops in unchanged") is therefore:

    keep the text of `class cchess_main` exactly as it is, and let every name it looks up resolve here.

`namespace()` returns that name table; `bind_reference_main(path)` reads a reference checkout's main.py, cuts out the
UNMODIFIED `class cchess_main` block and executes it over the table, returning the class.  Nothing of the reference is
vendored: the text is read from the user's checkout at run time (tests/test_gpu_reference_text.py does it with the copy
that oracle/stage_reference.py stages, and checks the result against the reference's own golden outputs).

A maintainer's patch to main.py is the same thing spelled statically: delete main.py:17-18 (the TensorFlow network
imports) and :23-1116, and write `from cchess_zero_b200.refapi import *` in their place."""
import copy
import os
import random
import time
from collections import defaultdict, deque, namedtuple

import numpy as np

from . import rules
from .mcts import MCTS_tree, leaf_node
from .net import policy_value_network, policy_value_network_gpus
from .rules import (GameBoard, c_PUCT, create_uci_labels, flipped_uci_labels, get_pieces_count, ind, is_kill_move,
                    labels_len, pieces_order, softmax, virtual_loss)

QueueItem = namedtuple("QueueItem", "feature future")   # main.py:229 (unused by the device search, kept for importers)
cut_off_depth = 30                                        # main.py:232

_TABLES = ("labels_array", "unflipped_index", "i2label", "label2i")
__all__ = ["GameBoard", "MCTS_tree", "leaf_node", "policy_value_network", "policy_value_network_gpus", "create_uci_labels",
           "flipped_uci_labels", "get_pieces_count", "is_kill_move", "softmax", "pieces_order", "ind", "labels_len", "c_PUCT",
           "virtual_loss", "cut_off_depth", "QueueItem", "np", "os", "random", "time", "copy", "deque", "defaultdict",
           "namedtuple"] + list(_TABLES)


def __getattr__(name):   # the label tables need the CUDA library; resolve them on first use
    if name in _TABLES:
        return getattr(rules, name)
    raise AttributeError(name)


def namespace():
    """dict of every module-level name `class cchess_main` (and ChessGame.py) reads from main.py."""
    ns = {k: globals()[k] for k in __all__ if k not in _TABLES}
    for k in _TABLES:
        ns[k] = getattr(rules, k)
    ns["flipped_labels"] = flipped_uci_labels(ns["labels_array"])   # main.py:213
    ns["__name__"] = "main"
    return ns


def class_source(main_py_text, name="cchess_main"):
    """The text of top-level `class <name>` in main.py, verbatim (up to the next top-level statement)."""
    lines = main_py_text.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("class %s(" % name) or l.startswith("class %s:" % name))
    end = len(lines)
    for i in range(start + 1, len(lines)):
        l = lines[i]
        if l and not l[0].isspace() and not l.startswith("#"):
            end = i
            break
    return "\n".join(lines[start:end]) + "\n", start + 1, end


def bind_reference_main(main_py_path, name="cchess_main", extra=None):
    """Executes the reference's own `class cchess_main` text over this package.  Returns (class, namespace)."""
    with open(main_py_path, encoding="utf-8") as f:
        text, first, _ = class_source(f.read(), name)
    ns = namespace()
    if extra:
        ns.update(extra)
    code = compile("\n" * (first - 1) + text, main_py_path, "exec")   # keep the reference's line numbers in tracebacks
    exec(code, ns)
    return ns[name], ns
