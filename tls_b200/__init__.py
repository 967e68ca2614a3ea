"""tls-b200: the TLS period/duration/T0 grid search on B200 (sm_100a) behind the
reference's ``transitleastsquares(t, y, dy).power(**kwargs)`` API.

Exports mirror ``/root/reference/transitleastsquares/__init__.py:13-18``
(``catalog_info`` needs astroquery + network and is out of scope)."""
from .main import transitleastsquares
from .helpers import cleaned_array, resample, transit_mask
from .grid import duration_grid, period_grid
from .stats import FAP, fold
from .results import transitleastsquaresresults
from .batch import batch_power, search_planets  # extras: batches and the iterative multi-planet recipe

__all__ = [
    "transitleastsquares", "cleaned_array", "resample", "transit_mask", "duration_grid",
    "period_grid", "FAP", "fold", "transitleastsquaresresults", "batch_power", "search_planets",
]
