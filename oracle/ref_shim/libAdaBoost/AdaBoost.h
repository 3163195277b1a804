// Stand-in -- TEST INFRASTRUCTURE.
#pragma once
class AdaBoostClassifier;
