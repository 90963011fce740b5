"""Import placeholder: plotting side effects of the reference become no-ops."""
