"""Drop-in overlay of the reference's `src.core` package (B200 / sm_100a implementations).

Modules this overlay does not provide (e.g. `src.core.text_corruptor`, used by
src/dnn_test_prio/case_study_imdb.py:12) fall through to the reference's own `src/core` when it
is on sys.path after this repository: the package path is extended over every `src/core`
directory of the `src` namespace package.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
