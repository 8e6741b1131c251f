from .tracer import Tracer, is_fx_tracing, symbolic_trace  # noqa: F401
