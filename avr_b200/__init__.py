"""avr_b200 -- B200-native (sm_100a) acoustic volume-rendering hot path of KMASAHIRO/AVR.

Public surface (mirrors the reference's ``renderer.py`` / ``model.py``):

* ``AVRRender(networks_fn, **render_cfg)``        <- renderer.py:14
* ``AVRModel(cfg)`` / ``AVRModel_complex(cfg)``   <- model.py:63 / :238
* ``Encoding`` / ``Network``                      <- tcnn.Encoding / tcnn.Network as used by model.py
* ``GradArena``                                   <- the DDP gradient all-reduce of avr_runner_ddp.py:98,257
* ``FusedAdam``                                   <- clip / NaN scrub / Adam of avr_runner.py:192-200
* ``Criterion(cfg_train, cfg_render)``            <- utils/criterion.py:7 (the training loss on rendered spectra)
* ``save_checkpoint`` / ``load_checkpoint``       <- avr_runner.py:104-154 (same file format, both directions)
* ``WaveLoader(base_folder, dataset_type, ...)``  <- datasets_loader.py:10 (the four on-disk dataset formats)

The compute path is hand-written CUDA behind the C-ABI of ``include/avr_b200.h``
(``avr_b200/libavr_b200.so``); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"

from .model import AVRModel, AVRModel_complex, Encoding, Network      # noqa: E402,F401
from .renderer import AVRRender                                       # noqa: E402,F401
from .ddp import GradArena                                            # noqa: E402,F401
from .optim import FusedAdam                                          # noqa: E402,F401
from .criterion import Criterion                                      # noqa: E402,F401
from .checkpoint import load_checkpoint, save_checkpoint, latest_checkpoint   # noqa: E402,F401
from .datasets import WaveLoader                                      # noqa: E402,F401
