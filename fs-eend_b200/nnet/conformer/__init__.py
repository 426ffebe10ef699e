"""Parameter containers of the LS-EEND Conformer-retention encoder (API mirror of LS-EEND/nnet/conformer/)."""
