"""Boundary glue kept from the reference's contract (xcolumns/utils.py): the "xcolumns" logger
gated by ``verbose`` (:21-44) and the ``__signature__`` rewriting that lets callers filter
kwargs by name (:209-230, used by experiments/utils.py:16-26)."""
import inspect
import logging
from typing import Callable, List, Optional

logging.basicConfig()
logger = logging.getLogger("xcolumns")
logger.setLevel(logging.INFO)


def log_info(msg: str, verbose: bool) -> None:
    if verbose:
        logger.info(msg)


def log_warning(msg: str, verbose: bool = True) -> None:
    if verbose:
        logger.warning(msg)


def add_kwargs_to_signature(func: Callable, func_with_kwargs: Callable, skip: Optional[List[str]] = None) -> Callable:
    """Expose on `func` (which takes **kwargs) the defaulted parameters of `func_with_kwargs`
    that it forwards to, minus `skip`, as its inspectable signature."""
    skip = set(skip or ())
    own = [p for p in inspect.signature(func).parameters.values() if p.kind is not inspect.Parameter.VAR_KEYWORD]
    forwarded = [p for p in inspect.signature(func_with_kwargs).parameters.values()
                 if p.default is not inspect.Parameter.empty and p.name not in skip
                 and p.name not in {q.name for q in own}]
    func.__signature__ = inspect.signature(func).replace(parameters=own + forwarded)
    return func
