"""ORACLE — test infrastructure only (see oracle/restatement.py). Never imported by metamorph_b200/."""
