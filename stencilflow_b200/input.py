from .base_node_class import Input  # noqa: F401  (module name kept for API parity)
