from .solve import FactorizedModel, biot_savart_film_to_film, factorize_model, solve, solve_batch
from .solve_film import (
    LinearSystem,
    TerminalSystems,
    factorize_linear_systems,
    solve_film,
    solve_for_terminal_current_stream,
)
from .utils import (
    FilmInfo,
    LambdaInfo,
    convert_field,
    field_conversion_factor,
)
