"""Scenario plugins (same file names as the reference's formation_gym/envs/)."""
