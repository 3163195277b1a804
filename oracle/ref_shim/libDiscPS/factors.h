// Stand-in -- TEST INFRASTRUCTURE.
#pragma once
