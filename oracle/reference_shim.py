"""Import the UNMODIFIED reference CPU renderer (this container only).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``/root/reference`` does not exist on the
GPU box, so nothing that runs there may call ``load_reference_renderer()``; tests that use it
skip when the tree is absent.  No reference source is copied: the module is executed from
where it lies.
"""
from __future__ import annotations

import importlib.util
import os

REFERENCE_ROOT = os.environ.get("AVR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "renderer_cpu.py"))


def load_reference_renderer():
    """-> the ``renderer_cpu`` module object (``AVRRender``, ``ray_directions`` ...)."""
    if not reference_available():
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")
    spec = importlib.util.spec_from_file_location(
        "avr_reference_renderer_cpu", os.path.join(REFERENCE_ROOT, "renderer_cpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
