"""Objects -> (dims, target indices) bridges (reference: photon_weave/state/utils/)."""
