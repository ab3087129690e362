"""`get_model(config)` -- same dispatch as the reference's models/epsnet/__init__.py:1-15."""


def get_model(config):
    if config.network == "dualenc":
        from .dualenc import DualEncoderEpsNetwork
        return DualEncoderEpsNetwork(config)
    elif config.network == "condensenc":
        from .condensenc import CondenseEncoderEpsNetwork
        return CondenseEncoderEpsNetwork(config)
    elif config.network == "dualenc_general":
        # the reference imports a module that does not exist in its tree (ImportError there)
        raise NotImplementedError("network 'dualenc_general' has no implementation in the reference")
    else:
        raise NotImplementedError("Unknown network: %s" % config.network)
