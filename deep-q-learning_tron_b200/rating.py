"""Batched policy-vs-policy rating: the reference's evaluation loops (play.py:72-98, ACKTR.py:409-421) play thousands of
`make_game(True, True, mode="fair", gamemode=...)` games one at a time through `Game.main_loop`; here all games advance in
lockstep on the GPU and each model is called once per tick on the whole batch."""
import torch

from .batch_env import BatchedTron


def _takes_extra(fn):
    """does fn(obs, extra) accept the side-feature argument?  Decided from the signature, never by catching TypeError."""
    import inspect
    try:
        params = list(inspect.signature(fn).parameters.values())
    except (TypeError, ValueError):
        return False
    if any(p.kind == p.VAR_POSITIONAL for p in params):
        return True
    return len([p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]) >= 2


@torch.no_grad()
def play_matches(model1, model2=None, n_games=10000, width=10, height=10, gamemode=None, slide=0.15, obs_enc="popup3",
                 obs_dtype=torch.float32, device="cuda", seed=0, max_ticks=None, spawn_mode="fair", layout="auto"):
    """Play n_games games of model1 (player 1) against model2 (player 2, default: model1).

    A model is anything with `.act(obs[n, P, W+2, H+2][, extra]) -> n actions` (like the reference's nets), a callable doing the
    same, or the string "minimax" for the batched scripted opponent (ACKTR.py:409-421 rates the agent against it).  Like
    Game.main_loop (tron/game.py:296-304), model1 receives extra = [[degree, weight_1]] per game ([n,2] f32, Game.get_multy(0)) and
    model2 receives extra = [[rate]] ([n,1] f32, Game.get_rate() = -((degree-30)*0.6)/100) when their act() takes a second argument.
    gamemode None | "ice" | "temper" (tron/game.py:163-178); spawns follow make_game(mode=spawn_mode).
    -> dict(p1_wins, p2_wins, draws, games, mean_ticks, p1_win_rating)   (p1_win_rating = p1/(p1+p2), play.py:94)
    """
    model2 = model2 or model1
    env = BatchedTron(n_games, width, height, device=device, obs_dtype=obs_dtype, obs_enc=obs_enc, auto_reset=False, seed=seed,
                      slide_mode=gamemode, slide_rate=slide, spawn_mode=spawn_mode, collect_stats=True, layout=layout, game_params=True)
    obs = env.reset()  # draws every game's degree / weights like Game.__init__ (game.py:83,87)
    act = torch.empty((n_games, 2), dtype=torch.uint8, device=env.device)
    limit = max_ticks or (width * height + 2)
    rate = (-((env.slide_params[:, 0].float() - 30) * 0.6) / 100).unsqueeze(1)  # Game.get_rate() (game.py:96-100); games never reset here

    def choose(m, o, player):
        if isinstance(m, str) and m == "minimax":  # the reference's scripted opponent, MinimaxPlayer(2, "voronoi")
            return env.minimax_actions(player)
        fn = m.act if hasattr(m, "act") else m
        if _takes_extra(fn):
            a = fn(o, env.extra[:, 0] if player == 1 else rate)
        else:
            a = fn(o)
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
