"""
``route_operation``: a state that lives inside an Envelope (integer index) or inside a
composite product state (tuple index) forwards the call to the method of the same name
on its container, passing itself (reference: photon_weave/state/utils/routing.py:8-47).
"""

from __future__ import annotations

from functools import wraps
from typing import Any, Callable


def route_operation() -> Callable:
    def decorator(method: Callable) -> Callable:
        name = "resize_fock" if method.__name__ == "resize" else method.__name__

        @wraps(method)
        def wrapper(self: Any, *args: Any, **kwargs: Any) -> Any:
            idx = self.index
            if isinstance(idx, int):
                holder, label = self.envelope, "Envelope"
            elif isinstance(idx, tuple):
                holder, label = self.composite_envelope, "CompositeEnvelope"
            else:
                return method(self, *args, **kwargs)  # the state is held by the object itself
            if not hasattr(holder, name):
                raise AttributeError(f"{label} has no method {name}")
            return getattr(holder, name)(*args, self, **kwargs)

        return wrapper

    return decorator
