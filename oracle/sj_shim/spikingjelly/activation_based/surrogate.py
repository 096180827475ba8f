from oracle.plif import ATan, Sigmoid, heaviside  # noqa: F401
