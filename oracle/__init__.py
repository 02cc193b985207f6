"""ORACLE -- CPU restatements of the reference hot path.  Test infrastructure only: the product package
(`mog_asr_b200`) must never import from here.  Parity unpinned (the reference has no tests)."""
