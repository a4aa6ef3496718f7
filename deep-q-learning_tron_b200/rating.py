"""Batched policy-vs-policy rating: the reference's evaluation loops (play.py:72-98, ACKTR.py:409-421) play thousands of
`make_game(True, True, mode="fair", gamemode=...)` games one at a time through `Game.main_loop`; here all games advance in
lockstep on the GPU and each model is called once per tick on the whole batch."""
import torch

from .batch_env import BatchedTron


@torch.no_grad()
def play_matches(model1, model2=None, n_games=10000, width=10, height=10, gamemode=None, slide=0.15, obs_enc="popup3",
                 obs_dtype=torch.float32, device="cuda", seed=0, max_ticks=None, spawn_mode="fair"):
    """Play n_games games of model1 (player 1) against model2 (player 2, default: model1).

    A model is anything with `.act(obs[n, P, W+2, H+2]) -> n actions` (like the reference's nets), a callable doing the same, or
    the string "minimax" for the batched scripted opponent (ACKTR.py:409-421 rates the agent against it).
    gamemode None | "ice" | "temper" (tron/game.py:163-178); spawns follow make_game(mode=spawn_mode).
    -> dict(p1_wins, p2_wins, draws, games, mean_ticks, p1_win_rating)   (p1_win_rating = p1/(p1+p2), play.py:94)
    """
    model2 = model2 or model1
    env = BatchedTron(n_games, width, height, device=device, obs_dtype=obs_dtype, obs_enc=obs_enc, auto_reset=False, seed=seed,
                      slide_mode=gamemode, slide_rate=slide, spawn_mode=spawn_mode, collect_stats=True)
    if gamemode == "temper":  # per-game degree / weights like Game.__init__ (game.py:83,87)
        g = torch.Generator(device="cpu").manual_seed(seed)
        env.slide_params[:, 0] = torch.randint(-30, 31, (n_games,), generator=g).to(torch.int8)
        env.slide_params[:, 1] = torch.randint(40, 102, (n_games,), generator=g).to(torch.int8)
        env.slide_params[:, 2] = torch.randint(40, 102, (n_games,), generator=g).to(torch.int8)
    obs = env.reset()
    act = torch.empty((n_games, 2), dtype=torch.uint8, device=env.device)
    limit = max_ticks or (width * height + 2)

    def choose(m, o, player):
        if isinstance(m, str) and m == "minimax":  # the reference's scripted opponent, MinimaxPlayer(2, "voronoi")
            return env.minimax_actions(player)
        a = m.act(o) if hasattr(m, "act") else m(o)
        return torch.as_tensor(a, device=env.device).reshape(-1).to(torch.uint8)

    ticks = 0
    while ticks < limit:
        act[:, 0] = choose(model1, obs[:, 0], 1)
        act[:, 1] = choose(model2, obs[:, 1], 2)
        res = env.step(act, obs=obs)
        ticks += 1
        if ticks % 4 == 0 and bool(res.done.all()):  # finished games stay frozen, so one flag read every few ticks suffices
            break
    st = env.stats_dict()
    ex = env.export()
    assert st["episodes"] == int(ex["done"].sum())
    decided = st["p1_wins"] + st["p2_wins"]
    return dict(p1_wins=st["p1_wins"], p2_wins=st["p2_wins"], draws=st["draws"], games=st["episodes"], unfinished=n_games - st["episodes"],
                mean_ticks=st["ep_ticks"] / max(1, st["episodes"]), p1_win_rating=(st["p1_wins"] / decided) if decided else float("nan"))
