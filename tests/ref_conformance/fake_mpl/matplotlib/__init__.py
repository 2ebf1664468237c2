"""TEST INFRASTRUCTURE ONLY: no-op stand-in so the reference's example scripts run headless."""
