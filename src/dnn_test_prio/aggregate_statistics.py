"""Drop-in overlay: `src.dnn_test_prio.aggregate_statistics` of the reference, served by simple_tip_b200."""
from simple_tip_b200.core.aggregate_statistics import *  # noqa: F401,F403
from simple_tip_b200.core.aggregate_statistics import AggStats, AggregateStatisticsCollector  # noqa: F401
