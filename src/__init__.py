"""`python -m src.run_modegpt ...` — the reference's module path, served by modegpt_b200.

The reference is run as `python -m src.run_modegpt` (README, tests.sh:87-133).  This package only
re-exports `modegpt_b200` under that name so those recipes keep working; all code lives there.
"""
import modegpt_b200 as _pkg

__path__ = _pkg.__path__
