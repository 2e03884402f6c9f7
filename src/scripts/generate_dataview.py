"""`python -m src.scripts.generate_dataview --inp=...` — same entry point as the reference."""
from lipreading_b200.dataview import generate_dataview  # noqa: F401


def main(argv=None):
    from lipreading_b200.cli import parseArgsForClassOrScript
    import inspect

    def generate_dataview_cli(inp="StephenColbert/nano2", vid_ext=".mp4", cap_ext=".vtt", out_ext=".npy",
                              timedelay=0, gen_vtx=False, force=False, seed=123456, gen_mouth=False):
        """ Generates dataviews for the given input directory of video/caption pairs. """
        return generate_dataview(inp, vid_ext, cap_ext, out_ext, timedelay, gen_vtx, force, seed, gen_mouth)
    args = vars(parseArgsForClassOrScript(generate_dataview_cli, argv))
    args.pop("verbosity", None)
    generate_dataview_cli(**args)


if __name__ == "__main__":
    main()
