"""Exception types of the state containers (reference: photon_weave/state/exceptions.py)."""


class NotExtractedException(Exception):
    pass


class EnvelopeAssignedException(Exception):
    pass


class EnvelopeAlreadyMeasuredException(Exception):
    pass


class MissingTemporalProfileArgumentException(Exception):
    pass
