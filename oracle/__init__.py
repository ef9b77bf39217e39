"""TEST INFRASTRUCTURE ONLY -- CPU restatement + reference runners.  Never imported by eesen_b200."""
