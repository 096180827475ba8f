from oracle.plif import OraclePLIF


class BaseNode(OraclePLIF):
    pass


class ParametricLIFNode(BaseNode):
    def __init__(self, init_tau=2.0, decay_input=True, v_threshold=1.0, v_reset=0.0,
                 surrogate_function=None, detach_reset=False, step_mode="s", backend="torch",
                 store_v_seq=False):
        super().__init__(init_tau, decay_input, v_threshold, v_reset, surrogate_function,
                         detach_reset, step_mode, backend, store_v_seq)


class LIFNode(BaseNode):
    pass
