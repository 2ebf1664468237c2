from .expression_interpreter import interpreter  # noqa: F401
