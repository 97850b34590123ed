from .base_node_class import Output  # noqa: F401  (module name kept for API parity)
