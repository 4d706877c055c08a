"""Integer enums of the game, numerically identical to the reference's ``game/enums.py:4-50``."""
from enum import IntEnum


class BuildingType(IntEnum):
    Settlement = 0
    City = 1


class PlayerId(IntEnum):
    White = 1
    Blue = 2
    Orange = 3
    Red = 4


class Resource(IntEnum):
    Empty = 0
    Brick = 1
    Wood = 2
    Ore = 3
    Sheep = 4
    Wheat = 5


class DevelopmentCard(IntEnum):
    Knight = 0
    VictoryPoint = 1
    YearOfPlenty = 2
    RoadBuilding = 3
    Monopoly = 4


class ActionTypes(IntEnum):
    PlaceSettlement = 0
    PlaceRoad = 1
    UpgradeToCity = 2
    BuyDevelopmentCard = 3
    PlayDevelopmentCard = 4
    ExchangeResource = 5
    ProposeTrade = 6
    RespondToOffer = 7
    MoveRobber = 8
    RollDice = 9
    EndTurn = 10
    StealResource = 11
    DiscardResource = 12
