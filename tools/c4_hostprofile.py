"""cProfile of the host side of config C4 (where the non-kernel time of a transient step goes)."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.config_runs as cr  # noqa: E402

pr = cProfile.Profile()
pr.enable()
cr.c4()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
