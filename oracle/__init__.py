"""CPU oracle for the EMAGE hot path - test infrastructure, never imported by the product."""
