"""TEST INFRASTRUCTURE ONLY — see oracle/README.md.  Never import from cmacionize_b200/."""
