"""CPU oracle of the BUFFER-X registration hot path -- TEST INFRASTRUCTURE, never imported by the product."""
