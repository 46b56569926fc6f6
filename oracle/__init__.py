"""ORACLE — test infrastructure (CPU restatement of the reference path).  See orc_common.h."""
