"""Overlay of the reference's `src.dnn_test_prio` package: only `aggregate_statistics` (the fit step of the
coverage criteria) is served from this repository; every other module (experiment drivers,
model handlers, case studies — out of scope) falls through to the reference's own directory."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
